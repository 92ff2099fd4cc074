// Continuous-world kernels (sm_100a):
//   K3 cw_rollout_kernel   fused random rollouts (mvmcts::RandomHeuristic over
//                          MvmctsState::Execute): lane = MCTS agent, floor(32/M) rollouts in
//                          flight per warp, finished rollouts are replaced at once from
//                          work the warps / CTAs claim at run time; agents that are only
//                          simulated live in shared-memory slots of the lanes
//   K4 cw_batch_kernel<EXECUTE>  one Execute for n states
//      cw_batch_kernel<RULES>    MvmctsState::EvaluateRules() for n states
//   K6 cw_batch_kernel<REPLAY>   replay of given action sequences with a full trace
//      cw_batch_kernel<QUERY>    CheckTerminal / GetNumActions / GetTerminalReward
// Every CTA stages the scenario blob it works on into shared memory with a TMA bulk copy
// (re-staged only when the scenario changes).
#include <algorithm>

#include "common.cuh"
#include "cw_device.cuh"
#include "philox.cuh"
#include "tma.cuh"

namespace brm {

namespace {

constexpr int kThreads = kCwThreads;
constexpr int kWarps = kThreads / 32;
constexpr double kFixScale = 16777216.0;  // 2^24: fixed-point scale of the per-leaf return sums

// shared memory layout: [blob_cap bytes blob][R*T q][R*T viol][Ps*kCwPasWords*T floats passive agents]
//                       (+ rollout kernel: [R*T u32 violation totals][kWarps*A leaf poses])
struct CwSmem {
  const CwBlob* B;
  uint8_t* q;     // this thread's column
  uint8_t* viol;  // this thread's column
  float* pas;     // this thread's column of passive-agent slots
};

__device__ __forceinline__ size_t cw_common_bytes(const CwCommon& c) {
  return static_cast<size_t>(c.blob_cap) + 2u * c.R * kThreads +
         static_cast<size_t>(c.Ps) * kCwPasWords * sizeof(float) * kThreads;
}

__device__ __forceinline__ CwSmem cw_smem(uint8_t* smem, const CwCommon& c) {
  CwSmem s;
  s.B = reinterpret_cast<const CwBlob*>(smem);
  s.q = smem + c.blob_cap + threadIdx.x;
  s.viol = smem + c.blob_cap + c.R * kThreads + threadIdx.x;
  s.pas = reinterpret_cast<float*>(smem + c.blob_cap + 2 * c.R * kThreads) + threadIdx.x;
  return s;
}

// (re)stage the blob of scenario `scn`; all threads of the CTA call this together
__device__ __forceinline__ void cw_stage(uint8_t* smem, uint64_t* bar, const CwCommon& c, int scn,
                                         int& cur_scn, uint32_t& phase) {
  if (scn == cur_scn) return;
  const bool first = cur_scn < 0;
  __syncthreads();  // everybody is done with the previous blob
  if (threadIdx.x == 0) {
    if (first) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t bytes = static_cast<uint32_t>(c.blob_bytes[scn]);
    mbar_expect_tx(bar, bytes);
    tma_load_1d(smem, c.blobs + static_cast<size_t>(scn) * c.blob_stride, bytes, bar);
  }
  if (first) __syncthreads();
  mbar_wait(bar, phase);
  phase ^= 1u;
  cur_scn = scn;
}

// lane -> (rollout group, MCTS agent)
struct CwMap {
  int lane, a, g, base;
  bool real;
};

__device__ __forceinline__ CwMap cw_map(const CwCommon& c) {
  CwMap m;
  m.lane = threadIdx.x & 31;
  m.a = m.lane % c.M;
  m.g = m.lane / c.M;
  m.base = m.g * c.M;
  m.real = m.g < c.G;
  return m;
}

__device__ __forceinline__ void cw_clear_agent(CwAgent& t) {
  t.x = t.y = t.th = t.v = 0.f;
  t.cs = 1.f;
  t.sn = 0.f;
  t.flags = 0;
  t.lane = -1;
  t.s = t.lat = 0.f;
  t.ex = t.ey = t.eth = t.ev = 0.f;
}

// agent `agent` of state i of batch S (no lane projection yet)
__device__ __forceinline__ void cw_load_agent(const brm_cw_states& S, int64_t i, int agent, bool on,
                                              CwAgent& t) {
  cw_clear_agent(t);
  if (on) {
    const float4 v = reinterpret_cast<const float4*>(S.agents)[static_cast<int64_t>(agent) * S.n + i];
    t.x = v.x; t.y = v.y; t.th = v.z; t.v = v.w;
    t.flags = S.flags[static_cast<int64_t>(agent) * S.n + i];
    sincosf(t.th, &t.sn, &t.cs);
  }
}

// whole state i for my lane: my MCTS agent into registers, my passive agents into their slots
__device__ __forceinline__ void cw_load_state(const brm_cw_states& S, const CwBlob& B, const CwCommon& c,
                                              const CwMap& m, const CwSmem& sm, int64_t i, bool on,
                                              CwAgent& t) {
  cw_load_agent(S, i, m.a, on, t);
  if (t.flags & BRM_CW_PRESENT) cw_frenet(B, t);
  for (int slot = 0; slot < c.Ps; ++slot) {
    const int j = c.M + slot * c.M + m.a;
    CwAgent p;
    cw_load_agent(S, i, j < c.A ? j : 0, on && j < c.A, p);
    if (p.flags & BRM_CW_PRESENT) cw_frenet(B, p);
    cw_pas_store(sm.pas, slot, p);
  }
}

__device__ __forceinline__ void cw_load_rules(const brm_cw_states& S, const CwCommon& c, int64_t i,
                                              int a, bool on, const CwSmem& sm, bool with_viol) {
  for (int sl = 0; sl < c.R; ++sl) {
    uint8_t q = 0, v = 0;
    if (on) {
      const int64_t o = (static_cast<int64_t>(a) * c.R + sl) * S.n + i;
      q = S.q[o];
      if (with_viol && S.viol) v = static_cast<uint8_t>(S.viol[o]);
    }
    sm.q[sl * kThreads] = q;
    sm.viol[sl * kThreads] = v;
  }
}

// Predict for my passive agents: they follow primitive 0 of their own table
// (BehaviorConstantAcceleration in the reference's prediction settings)
template <bool GRID = false>
__device__ __forceinline__ uint32_t cw_move_passive(const CwBlob& B, const CwCommon& c, const CwMap& m,
                                                    const CwSmem& sm, bool live) {
  uint32_t moved = 0;
  for (int slot = 0; slot < c.Ps; ++slot) {
    const int j = c.M + slot * c.M + m.a;
    if (!live || j >= c.A) continue;
    CwAgent p;
    cw_pas_load(sm.pas, slot, p);
    if ((p.flags & BRM_CW_PRESENT) && (p.flags & BRM_CW_VALID)) {
      cw_integrate(B, p, j, 0u);
      cw_frenet<GRID>(B, p);
      cw_pas_store(sm.pas, slot, p);
      ++moved;
    }
  }
  return moved;
}

// ------------------------------------------------------------------ K3 rollout
struct CwRolloutArgs {
  CwCommon c;
  brm_cw_states leaves;
  brm_cw_rollout_out out;
  long long* returns_fx;             // [leaves][M*V] fixed-point sums, zeroed before the launch
  unsigned long long* work_counter;  // next unclaimed rollout (one scenario) / unit (several), zeroed
  uint64_t seed, rollout_base;
  int32_t rollouts_per_leaf, max_steps, first_exp;
  int32_t single_scenario;  // 1: every leaf lives in scenario 0 -> warps draw from the flat rollout range
  int32_t chunk;            // rollouts a warp claims at a time (multiple of G)
  int32_t parts;            // several scenarios: a leaf's rollouts are handed out in `parts` units
  float gamma;
};

// the pose of a leaf's agents, kept per warp so that a finished rollout can be replaced
// without going back to global memory or redoing the lane projection
struct CwPose {
  float x, y, th, v, cs, sn, s, lat;
  int lane;
  uint32_t flags;
};

__device__ __forceinline__ void cw_from_pose(const CwPose& p, CwAgent& t) {
  t.x = p.x; t.y = p.y; t.th = p.th; t.v = p.v; t.cs = p.cs; t.sn = p.sn;
  t.s = p.s; t.lat = p.lat; t.lane = p.lane; t.flags = p.flags;
  t.ex = t.ey = t.eth = t.ev = 0.f;
}

// What a lane has accumulated for the leaf its recent rollouts belonged to.  Returns are
// converted to fixed point per rollout, so the integer sums do not depend on which warp ran
// which rollout or in which order (bitwise reproducible under dynamic scheduling).
struct CwLaneAcc {
  long long fx[BRM_MAX_REWARD];
  int leaf;
  uint32_t steps;
};

__device__ __forceinline__ void cw_acc_flush(const CwRolloutArgs& args, CwLaneAcc& acc, uint32_t* viol_tot, int a) {
  const int M = args.c.M, V = args.c.V, R = args.c.R;
  if (acc.leaf >= 0) {
#pragma unroll
    for (int v = 0; v < BRM_MAX_REWARD; ++v)
      if (v < V && acc.fx[v])
        atomicAdd(reinterpret_cast<unsigned long long*>(&args.returns_fx[(static_cast<int64_t>(acc.leaf) * M + a) * V + v]),
                  static_cast<unsigned long long>(acc.fx[v]));
    if (args.out.viol_sum) {
      for (int sl = 0; sl < R; ++sl) {
        const uint32_t x = viol_tot[sl * kThreads];
        if (x)
          atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.viol_sum[(static_cast<int64_t>(acc.leaf) * M + a) * R + sl]),
                    static_cast<unsigned long long>(x));
      }
    }
    if (args.out.agent_steps && acc.steps)
      atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.agent_steps[acc.leaf]),
                static_cast<unsigned long long>(acc.steps));
  }
#pragma unroll
  for (int v = 0; v < BRM_MAX_REWARD; ++v) acc.fx[v] = 0;
  for (int sl = 0; sl < R; ++sl) viol_tot[sl * kThreads] = 0u;
  acc.steps = 0;
}

// One warp runs rollouts until `fetch` has no more work and all its lane groups are done.
// Up to G rollouts are in flight (one per lane group); when one ends its group takes the
// next rollout id of the warp's current range at once, and an empty range is replaced by
// `fetch(lo, hi)` (warp-uniform; a range [lo, hi) of flat rollout ids inside ONE leaf, or
// lo >= hi when the work is exhausted).  Lanes therefore only idle at the very end.
template <bool PASSIVE, bool GRID, class Fetch>
__device__ __forceinline__ void cw_run_warp(const CwRolloutArgs& args, const CwBlob& B, const CwSmem& sm,
                                            uint32_t* viol_tot, CwPose* pose_w, CwLaneAcc& lacc, Fetch fetch) {
  const CwCommon& c = args.c;
  const CwMap m = cw_map(c);
  const int lane = m.lane, a = m.a, base = m.base;
  const bool real = m.real;
  const int A = c.A, M = c.M, V = c.V, R = c.R;
  const uint32_t full = 0xffffffffu;
  const int n = args.rollouts_per_leaf;

  CwAgent t;
  cw_clear_agent(t);
  float acc[BRM_MAX_REWARD] = {0.f, 0.f, 0.f, 0.f};
  float disc = 1.f;
  int horizon = 0, k = 0;
  int flat = 0, my_leaf = -1;  // the rollout my group runs (flat id < 2^31, checked by the launcher)
  uint32_t steps = 0;
  bool busy = false, done = false;
  Philox4 blk;
  blk.w[0] = blk.w[1] = blk.w[2] = blk.w[3] = 0;
  // warp-uniform: unclaimed part of the current range, the leaf it lies in and that leaf's data
  int lo = 0, hi = 0, leaf = -1;
  int horizon0 = 0;
  bool leaf_terminal = false, exhausted = false;

  for (;;) {
    // ---- hand new rollouts to idle groups (ego lanes vote, ranks give distinct ids)
    for (;;) {
      const uint32_t idle = __ballot_sync(full, real && !busy && a == 0);
      if (!idle || exhausted) break;
      if (lo >= hi) {
        fetch(lo, hi);
        if (lo >= hi) {
          exhausted = true;
          break;
        }
        const int lf = lo / n;
        if (lf != leaf) {
          // leaf pose of every agent -> per-warp shared copy (lane l < A loads and projects agent l)
          leaf = lf;
          CwAgent t0;
          cw_load_agent(args.leaves, leaf, lane < A ? lane : 0, lane < A, t0);
          if (t0.flags & BRM_CW_PRESENT) cw_frenet<GRID>(B, t0);
          __syncwarp();
          if (lane < A) {
            CwPose p;
            p.x = t0.x; p.y = t0.y; p.th = t0.th; p.v = t0.v; p.cs = t0.cs; p.sn = t0.sn;
            p.s = t0.s; p.lat = t0.lat; p.lane = t0.lane; p.flags = t0.flags;
            pose_w[lane] = p;
          }
          __syncwarp();
          horizon0 = args.leaves.horizon[leaf];
          leaf_terminal = args.leaves.terminal[leaf] != 0;
        }
      }
      const int cand = lo + __popc(idle & ((1u << base) - 1u));
      if (real && !busy && cand < hi) {
        flat = cand;
        my_leaf = leaf;
        cw_from_pose(pose_w[a], t);
        if (PASSIVE) {
          for (int slot = 0; slot < c.Ps; ++slot) {
            const int j = M + slot * M + a;
            CwAgent p;
            cw_clear_agent(p);
            if (j < A) cw_from_pose(pose_w[j], p);
            cw_pas_store(sm.pas, slot, p);
          }
        }
        for (int sl = 0; sl < R; ++sl) {
          sm.q[sl * kThreads] = args.leaves.q[(static_cast<int64_t>(a) * R + sl) * args.leaves.n + leaf];
          sm.viol[sl * kThreads] = 0;
        }
#pragma unroll
        for (int v = 0; v < BRM_MAX_REWARD; ++v) acc[v] = 0.f;
        disc = args.first_exp ? args.gamma : 1.f;
        horizon = horizon0;
        k = 0;
        steps = 0;
        busy = true;
        done = leaf_terminal;
      }
      lo = min(hi, lo + __popc(idle));
    }
    if (!__any_sync(full, busy)) break;

    // ---- one Execute for every busy group
    const uint64_t rid = args.rollout_base + static_cast<uint64_t>(flat);
    const bool live = busy && !done && k < args.max_steps;
    const bool movable = (t.flags & BRM_CW_PRESENT) && (t.flags & BRM_CW_VALID);
    uint32_t act = 0;
    if (live) {
      if ((k & 3) == 0) blk = action_block(args.seed, rid, k, a);
      act = draw_action(pick_word(blk, k), movable ? B.n_prims[a] : 1u);
    }
    const float in_v = t.v, in_th = t.th;
    const uint32_t in_flags = t.flags;
    if (live && movable) {
      cw_integrate(B, t, a, act);
      cw_frenet<GRID>(B, t);
      ++steps;
    }
    if (PASSIVE) steps += cw_move_passive<GRID>(B, c, m, sm, live);
    CwStepOut o;
    cw_evaluate<false, PASSIVE>(B, t, a, base, real, live, in_v, in_th, in_flags, horizon - 1, sm.q, sm.viol,
                                kThreads, sm.pas, o);
    if (live) {
#pragma unroll
      for (int v = 0; v < BRM_MAX_REWARD; ++v) acc[v] += disc * o.r[v];
      disc *= args.gamma;
      --horizon;
      ++k;
      if (o.terminal) done = true;
    }
    // ---- rollouts that just ended: terminal reward, totals, optional per-rollout outputs
    if (busy && (done || k >= args.max_steps)) {
      if (my_leaf != lacc.leaf) {
        cw_acc_flush(args, lacc, viol_tot, a);
        lacc.leaf = my_leaf;
      }
      float tr[BRM_MAX_REWARD];
      cw_terminal_reward(B, t, a, true, sm.q, kThreads, tr);
#pragma unroll
      for (int v = 0; v < BRM_MAX_REWARD; ++v) {
        acc[v] += disc * tr[v];
        lacc.fx[v] += __double2ll_rn(static_cast<double>(acc[v]) * kFixScale);
      }
      lacc.steps += steps;
      const uint64_t local = static_cast<uint64_t>(flat);
      if (args.out.ro_returns)
#pragma unroll
        for (int v = 0; v < BRM_MAX_REWARD; ++v)
          if (v < V) args.out.ro_returns[(local * M + a) * V + v] = acc[v];
      for (int sl = 0; sl < R; ++sl) {
        const uint8_t vv = sm.viol[sl * kThreads];
        viol_tot[sl * kThreads] += vv;
        if (args.out.ro_viol) args.out.ro_viol[(local * M + a) * R + sl] = vv;
        if (args.out.ro_q) args.out.ro_q[(local * M + a) * R + sl] = sm.q[sl * kThreads];
      }
      if (args.out.ro_agents)
        reinterpret_cast<float4*>(args.out.ro_agents)[local * A + a] = make_float4(t.x, t.y, t.th, t.v);
      if (args.out.ro_flags) args.out.ro_flags[local * A + a] = static_cast<uint8_t>(t.flags);
      if (PASSIVE && (args.out.ro_agents || args.out.ro_flags)) {
        for (int slot = 0; slot < c.Ps; ++slot) {
          const int j = M + slot * M + a;
          if (j >= A) continue;
          CwAgent p;
          cw_pas_load(sm.pas, slot, p);
          if (args.out.ro_agents)
            reinterpret_cast<float4*>(args.out.ro_agents)[local * A + j] = make_float4(p.x, p.y, p.th, p.v);
          if (args.out.ro_flags) args.out.ro_flags[local * A + j] = static_cast<uint8_t>(p.flags);
        }
      }
      if (args.out.ro_steps && a == 0) args.out.ro_steps[local] = k;
      busy = false;
      t.flags = 0;  // an idle group must not look present to anybody
    }
  }
  __syncwarp();
}

template <bool PASSIVE, bool GRID>
__global__ void __launch_bounds__(kThreads, kCwRolloutCtasPerSm) cw_rollout_kernel(const CwRolloutArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned long long s_unit;  // several scenarios: the unit the CTA works on ...
  __shared__ int s_next, s_end;          // ... and the unclaimed part of its rollouts (leaf-local ids)
  const CwCommon& c = args.c;
  const CwSmem sm = cw_smem(smem, c);
  const CwBlob& B = *sm.B;
  const int widx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* extra = smem + cw_common_bytes(c);
  uint32_t* viol_tot = reinterpret_cast<uint32_t*>(extra) + threadIdx.x;
  CwPose* pose_w = reinterpret_cast<CwPose*>(extra + 4 * c.R * kThreads) + widx * c.A;
  int cur_scn = -1;
  uint32_t phase = 0;
  const int n = args.rollouts_per_leaf;
  const int a = lane % c.M;
  CwLaneAcc lacc;
  lacc.leaf = -1;
  cw_acc_flush(args, lacc, viol_tot, a);  // zero
  if (args.single_scenario) {
    // one scenario: warps claim chunks of the flat rollout range [0, leaves*n) from a global
    // counter as they run dry (rollout lengths differ a lot between leaves, a static split
    // leaves warps idle at the end)
    cw_stage(smem, &bar, c, 0, cur_scn, phase);
    const int N = static_cast<int>(args.leaves.n) * n;
    int c_lo = 0, c_hi = 0;  // claimed, not yet handed to the lane groups
    auto fetch = [&](int& lo, int& hi) {
      if (c_lo >= c_hi) {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(args.work_counter, static_cast<unsigned long long>(args.chunk));
        v = __shfl_sync(0xffffffffu, v, 0);
        c_lo = static_cast<int>(min(v, static_cast<unsigned long long>(N)));
        c_hi = min(c_lo + args.chunk, N);
      }
      // hand out the part that lies in one leaf
      lo = c_lo;
      hi = min(c_hi, (c_lo / n + 1) * n);
      c_lo = hi;
    };
    cw_run_warp<PASSIVE, GRID>(args, B, sm, viol_tot, pose_w, lacc, fetch);
  } else {
    // several scenarios: the blob is CTA-wide, so a CTA works on one leaf at a time.  Units
    // (a leaf's rollouts, or one of `parts` slices of them) are claimed from a global counter;
    // inside a unit the warps draw G rollouts at a time from a shared-memory counter.
    const int n_units = static_cast<int>(args.leaves.n) * args.parts;
    for (;;) {
      __syncthreads();  // everybody is done with the previous unit's counters
      if (threadIdx.x == 0) {
        const unsigned long long u = atomicAdd(args.work_counter, 1ull);
        s_unit = u;
        if (u < static_cast<unsigned long long>(n_units)) {
          const int part = static_cast<int>(u % args.parts);
          s_next = static_cast<int>(static_cast<int64_t>(n) * part / args.parts);
          s_end = static_cast<int>(static_cast<int64_t>(n) * (part + 1) / args.parts);
        }
      }
      __syncthreads();
      if (s_unit >= static_cast<unsigned long long>(n_units)) break;
      const int leaf = static_cast<int>(s_unit) / args.parts;
      cw_stage(smem, &bar, c, args.leaves.scenario[leaf], cur_scn, phase);
      const int end = s_end;
      auto fetch = [&](int& lo, int& hi) {
        int v = 0;
        if (lane == 0) v = atomicAdd(&s_next, c.G);
        v = __shfl_sync(0xffffffffu, v, 0);
        const int b = min(v, end), e = min(v + c.G, end);
        lo = leaf * n + b;
        hi = leaf * n + e;
      };
      cw_run_warp<PASSIVE, GRID>(args, B, sm, viol_tot, pose_w, lacc, fetch);
    }
  }
  cw_acc_flush(args, lacc, viol_tot, a);
}

// fixed point -> double, and the agents' sums mixed: value_i = own_i + coop * sum_{j != i} own_j
// (mvmcts COOP_FACTOR, src/util.cpp:28-29; 0 in every test of the reference: unpinned)
__global__ void cw_finalize_returns_kernel(const long long* __restrict__ fx, double* __restrict__ out,
                                           int64_t count, int M, int V, double coop) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double own = static_cast<double>(fx[i]) * (1.0 / kFixScale);
  double others = 0.0;
  if (coop != 0.0) {
    const int64_t leaf = i / (static_cast<int64_t>(M) * V);
    const int v = static_cast<int>(i % V);
    for (int j = 0; j < M; ++j) {
      const int64_t o = (leaf * M + j) * V + v;
      if (o != i) others += static_cast<double>(fx[o]) * (1.0 / kFixScale);
    }
  }
  out[i] = coop != 0.0 ? own + coop * others : own;
}

// ------------------------------------------------------------------ K4 / K6 / queries
// states are mapped G per warp, lane = MCTS agent
struct CwBatchArgs {
  CwCommon c;
  brm_cw_states in, out;
  const uint8_t* actions;  // execute: [M][n]; replay: [B][T][M]
  float* rewards;          // execute: [M][V][n]
  // queries
  uint8_t* num_actions;    // [M][n]
  float* term_reward;      // [M][V][n]
  int32_t do_check_terminal;
  // replay
  int32_t T;
  brm_cw_trace tr;
  int64_t warp_units;  // ceil(n / G)
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) cw_batch_kernel(const CwBatchArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const CwCommon& c = args.c;
  const CwSmem sm = cw_smem(smem, c);
  const CwBlob& B = *sm.B;
  const CwMap m = cw_map(c);
  const int widx = threadIdx.x >> 5;
  const int A = c.A, M = c.M, V = c.V, R = c.R, G = c.G;
  const int a = m.a, g = m.g, base = m.base;
  const bool real = m.real;
  const int64_t n = args.in.n;
  int cur_scn = -1;
  uint32_t phase = 0;
  // A CTA walks over blocks of kWarps*G consecutive states; the batch is homogeneous in
  // scenario (checked by the host API), so the blob is staged once.
  const int64_t cta_units = (args.warp_units + kWarps - 1) / kWarps;
  for (int64_t cu = blockIdx.x; cu < cta_units; cu += gridDim.x) {
    const int64_t first_state = cu * kWarps * G;
    cw_stage(smem, &bar, c, args.in.scenario[first_state < n ? first_state : n - 1], cur_scn, phase);
    const int64_t i = (cu * kWarps + widx) * G + g;
    const bool on = real && i < n;
    const int64_t ii = on ? i : 0;
    CwAgent t;
    cw_load_state(args.in, B, c, m, sm, ii, on, t);
    cw_load_rules(args.in, c, ii, a, on, sm, true);
    int horizon = on ? args.in.horizon[ii] : 0;

    if (MODE == MODE_QUERY) {
      if (args.do_check_terminal) {
        // CheckTerminal (:266-285) for the ego of each state
        float cx[4], cy[4];
        cw_corners(B, t, a, cx, cy);
        const bool present = on && (t.flags & BRM_CW_PRESENT);
        bool coll = false;
        __syncwarp();
        for (int j = 0; j < A; ++j) {
          const CwPeer p = j < M ? cw_fetch(t, base + j) : cw_fetch_passive(B, sm.pas, j, a, real);
          if (j == a || !(p.flags & BRM_CW_PRESENT) || !present) continue;
          coll = cw_rect_overlap(B, t, a, p, j) || coll;
        }
        if (on && a == 0) {
          const bool term = !present || horizon == 0 || cw_out_of_map(B, t, a, cx, cy) || coll ||
                            cw_at_goal(B, t, a, cx, cy);
          args.in.terminal[i] = term ? 1 : 0;
        }
      }
      if (on) {
        if (args.num_actions) {
          const bool movable = (t.flags & BRM_CW_PRESENT) && (t.flags & BRM_CW_VALID);
          args.num_actions[static_cast<int64_t>(a) * n + i] = movable ? B.n_prims[a] : 1;
        }
        if (args.term_reward) {
          float tr[BRM_MAX_REWARD];
          cw_terminal_reward(B, t, a, true, sm.q, kThreads, tr);
          for (int v = 0; v < V; ++v) args.term_reward[(static_cast<int64_t>(a) * V + v) * n + i] = tr[v];
        }
      }
      continue;
    }

    if (MODE == MODE_EXECUTE || MODE == MODE_RULES) {
      // MODE_RULES: EvaluateRules() on the state as it is (no kinematic step, horizon and
      // terminal flag untouched)
      constexpr bool kRules = MODE == MODE_RULES;
      const bool movable = (t.flags & BRM_CW_PRESENT) && (t.flags & BRM_CW_VALID);
      uint32_t act = 0;
      if (on && !kRules) act = args.actions[static_cast<int64_t>(a) * n + i];
      if (!movable) act = 0;
      const float in_v = t.v, in_th = t.th;
      const uint32_t in_flags = t.flags;
      if (on && movable && !kRules) {
        cw_integrate(B, t, a, act);
        cw_frenet(B, t);
      }
      if (!kRules) cw_move_passive(B, c, m, sm, on);
      CwStepOut o;
      cw_evaluate<false, true>(B, t, a, base, real, on, in_v, in_th, in_flags, horizon - 1, sm.q, sm.viol,
                               kThreads, sm.pas, o, kRules);
      if (on) {
        reinterpret_cast<float4*>(args.out.agents)[static_cast<int64_t>(a) * n + i] =
            make_float4(t.x, t.y, t.th, t.v);
        args.out.flags[static_cast<int64_t>(a) * n + i] = static_cast<uint8_t>(t.flags);
        for (int slot = 0; slot < c.Ps; ++slot) {
          const int j = M + slot * M + a;
          if (j >= A) continue;
          CwAgent p;
          cw_pas_load(sm.pas, slot, p);
          reinterpret_cast<float4*>(args.out.agents)[static_cast<int64_t>(j) * n + i] =
              make_float4(p.x, p.y, p.th, p.v);
          args.out.flags[static_cast<int64_t>(j) * n + i] = static_cast<uint8_t>(p.flags);
        }
        if (a == 0) {
          args.out.horizon[i] = kRules ? horizon : horizon - 1;
          args.out.terminal[i] = kRules ? args.in.terminal[i] : (o.terminal ? 1 : 0);
          args.out.scenario[i] = args.in.scenario[i];
        }
        for (int v = 0; v < V; ++v) args.rewards[(static_cast<int64_t>(a) * V + v) * n + i] = o.r[v];
        for (int sl = 0; sl < R; ++sl) {
          const int64_t off = (static_cast<int64_t>(a) * R + sl) * n + i;
          args.out.q[off] = sm.q[sl * kThreads];
          if (args.out.viol) {
            const uint32_t prev = args.in.viol ? args.in.viol[off] : 0u;
            const uint32_t inc = static_cast<uint8_t>(sm.viol[sl * kThreads] - static_cast<uint8_t>(prev));
            args.out.viol[off] = prev + inc;
          }
        }
      }
      continue;
    }

    if (MODE == MODE_REPLAY) {
      const int T = args.T;
      bool done = !on || args.in.terminal[ii] != 0;
      int steps = 0;
      uint32_t viol32[BRM_CW_MAX_SLOTS];
      for (int sl = 0; sl < BRM_CW_MAX_SLOTS; ++sl) viol32[sl] = 0;
      for (int sl = 0; sl < R; ++sl) sm.viol[sl * kThreads] = 0;
      for (int k = 0; k < T; ++k) {
        const bool live = !done;
        if (!__any_sync(0xffffffffu, live)) break;
        const int64_t bt = ii * T + k;
        const bool movable = (t.flags & BRM_CW_PRESENT) && (t.flags & BRM_CW_VALID);
        uint32_t act = 0;
        if (live && movable) act = args.actions[bt * M + a];
        const float in_v = t.v, in_th = t.th;
        const uint32_t in_flags = t.flags;
        if (live && movable) {
          cw_integrate(B, t, a, act);
          cw_frenet(B, t);
        }
        cw_move_passive(B, c, m, sm, live);
        CwStepOut o;
        cw_evaluate<true, true>(B, t, a, base, real, live, in_v, in_th, in_flags, horizon - 1, sm.q, sm.viol,
                          kThreads, sm.pas, o);
        if (live) {
          --horizon;
          ++steps;
          reinterpret_cast<float4*>(args.tr.agents_out)[bt * A + a] = make_float4(t.x, t.y, t.th, t.v);
          args.tr.flags_out[bt * A + a] = static_cast<uint8_t>(t.flags);
          float* fo = args.tr.frenet_out + (bt * A + a) * 3;
          fo[0] = static_cast<float>(t.lane);
          fo[1] = t.s;
          fo[2] = t.lat;
          for (int slot = 0; slot < c.Ps; ++slot) {
            const int j = M + slot * M + a;
            if (j >= A) continue;
            CwAgent p;
            cw_pas_load(sm.pas, slot, p);
            reinterpret_cast<float4*>(args.tr.agents_out)[bt * A + j] = make_float4(p.x, p.y, p.th, p.v);
            args.tr.flags_out[bt * A + j] = static_cast<uint8_t>(p.flags);
            float* fp = args.tr.frenet_out + (bt * A + j) * 3;
            fp[0] = static_cast<float>(p.lane);
            fp[1] = p.s;
            fp[2] = p.lat;
          }
          if (a == 0) args.tr.terminal_out[bt] = o.terminal ? 1 : 0;
          for (int v = 0; v < V; ++v) args.tr.rewards_out[(bt * M + a) * V + v] = o.r[v];
          for (int w = 0; w < 4; ++w) args.tr.labels_out[(bt * M + a) * 4 + w] = o.w[w];
          for (int sl = 0; sl < R; ++sl) {
            viol32[sl] += sm.viol[sl * kThreads];
            sm.viol[sl * kThreads] = 0;
            args.tr.q_out[(bt * M + a) * R + sl] = sm.q[sl * kThreads];
            args.tr.viol_out[(bt * M + a) * R + sl] = viol32[sl];
          }
          if (o.terminal) done = true;
        }
      }
      if (on) {
        if (a == 0) args.tr.steps_out[i] = steps;
        float tr[BRM_MAX_REWARD];
        cw_terminal_reward(B, t, a, true, sm.q, kThreads, tr);
        for (int v = 0; v < V; ++v) args.tr.term_reward_out[(i * M + a) * V + v] = tr[v];
      }
    }
  }
}

}  // namespace

size_t cw_rollout_smem_bytes(const CwLaunchCtx& ctx) {
  return ctx.smem_bytes + 4u * static_cast<size_t>(ctx.c.R) * kThreads +
         static_cast<size_t>(kWarps) * ctx.c.A * sizeof(CwPose);
}

int cw_launch_rollout(brm_engine* e, const CwLaunchCtx& ctx, int n_scenarios, const brm_cw_states& leaves,
                      const brm_rollout_params& rp, const brm_cw_rollout_out& out) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cw_rollout_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cw_rollout_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cw_rollout_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cw_rollout_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int M = ctx.c.M, V = ctx.c.V, R = ctx.c.R;
  CwRolloutArgs a;
  a.c = ctx.c;
  a.leaves = leaves;
  a.out = out;
  a.seed = rp.seed;
  a.rollout_base = rp.rollout_base;
  a.rollouts_per_leaf = rp.rollouts_per_leaf;
  a.max_steps = rp.max_steps;
  a.first_exp = rp.first_discount_exp;
  a.gamma = static_cast<float>(rp.discount_factor);
  a.single_scenario = n_scenarios == 1 ? 1 : 0;
  const size_t fx_count = static_cast<size_t>(leaves.n) * M * V;
  int rc = e->ensure_ws((fx_count + 1) * sizeof(long long));
  if (rc != BRM_OK) return rc;
  a.returns_fx = static_cast<long long*>(e->d_ws);
  a.work_counter = reinterpret_cast<unsigned long long*>(a.returns_fx + fx_count);
  BRM_CUDA_CHECK(cudaMemsetAsync(a.returns_fx, 0, (fx_count + 1) * sizeof(long long), e->stream));
  if (out.viol_sum)
    BRM_CUDA_CHECK(cudaMemsetAsync(out.viol_sum, 0, static_cast<size_t>(leaves.n) * M * R * sizeof(uint64_t),
                                   e->stream));
  if (out.agent_steps)
    BRM_CUDA_CHECK(cudaMemsetAsync(out.agent_steps, 0, static_cast<size_t>(leaves.n) * sizeof(uint64_t),
                                   e->stream));
  // persistent grid: at most the number of CTAs that are resident at once (SMs x occupancy);
  // the CTAs claim their work from a counter
  const size_t smem = cw_rollout_smem_bytes(ctx);
  int per_sm = 0;
  // instances of the kernel: agents outside the MCTS list (kept in shared-memory slots) or not,
  // candidate blocks of the lane projection from the grid or from a pass over all boxes
  const bool passive = ctx.c.Ps > 0;
  auto kernel = passive ? (ctx.c.use_grid ? cw_rollout_kernel<true, true> : cw_rollout_kernel<true, false>)
                        : (ctx.c.use_grid ? cw_rollout_kernel<false, true> : cw_rollout_kernel<false, false>);
  BRM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
  int64_t grid = static_cast<int64_t>(e->sm_count) * (per_sm > 0 ? per_sm : 1);
  const int64_t total = static_cast<int64_t>(leaves.n) * rp.rollouts_per_leaf;
  BRM_REQUIRE(total + 4 * 32 < (1ll << 31), "brm_cw_rollout: %lld rollouts in one call (limit 2^31)",
              static_cast<long long>(total));
  const int64_t round = static_cast<int64_t>(kWarps) * ctx.c.G;  // rollouts a CTA has in flight
  a.chunk = ctx.c.G;
  a.parts = 1;
  if (a.single_scenario) {
    // at least one rollout per lane group, otherwise shrink the grid
    const int64_t need = (total + round - 1) / round;
    if (grid > need) grid = need;
    // a warp claims 1..4 rounds of rollouts at a time, about a quarter of its share
    const int64_t share = total / (grid * kWarps * ctx.c.G);
    a.chunk = ctx.c.G * static_cast<int>(std::min<int64_t>(4, std::max<int64_t>(1, share / 4)));
  } else {
    // cut the leaves into slices (at least one round of the CTA each) until there are about
    // two units per resident CTA to balance with (measured on the W workload: 1 / 2 / 3 / 4
    // slices per leaf -> 1.46 / 1.89 / 2.02 / 1.89 G agent-steps/s)
    while (static_cast<int64_t>(leaves.n) * a.parts < 2 * grid &&
           rp.rollouts_per_leaf / (a.parts + 1) >= round)
      ++a.parts;
    const int64_t units = static_cast<int64_t>(leaves.n) * a.parts;
    if (grid > units) grid = units;
  }
  if (grid < 1) grid = 1;
  e->time_begin();
  kernel<<<static_cast<int>(grid), kThreads, smem, e->stream>>>(a);
  e->launches++;
  cw_finalize_returns_kernel<<<static_cast<int>((fx_count + 255) / 256), 256, 0, e->stream>>>(
      a.returns_fx, out.returns_sum, static_cast<int64_t>(fx_count), M, V, rp.coop_factor);
  e->launches++;
  e->time_end();
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

static int64_t batch_grid(const brm_engine* e, int64_t warp_units) {
  int64_t grid = (warp_units + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(e->sm_count) * 4;
  if (grid > cap) grid = cap;
  return grid < 1 ? 1 : grid;
}

int cw_launch_batch(brm_engine* e, const CwLaunchCtx& ctx, int mode, const brm_cw_states& in,
                    const brm_cw_states* out, const uint8_t* actions, float* rewards,
                    uint8_t* num_actions, float* term_reward, int do_check_terminal, int32_t T,
                    const brm_cw_trace* tr) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cw_batch_kernel<MODE_EXECUTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cw_batch_kernel<MODE_QUERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cw_batch_kernel<MODE_REPLAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cw_batch_kernel<MODE_RULES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  CwBatchArgs a;
  memset(&a, 0, sizeof(a));
  a.c = ctx.c;
  a.in = in;
  if (out) a.out = *out;
  a.actions = actions;
  a.rewards = rewards;
  a.num_actions = num_actions;
  a.term_reward = term_reward;
  a.do_check_terminal = do_check_terminal;
  a.T = T;
  if (tr) a.tr = *tr;
  a.warp_units = (static_cast<int64_t>(in.n) + ctx.c.G - 1) / ctx.c.G;
  const int grid = static_cast<int>(batch_grid(e, a.warp_units));
  if (mode == MODE_EXECUTE)
    cw_batch_kernel<MODE_EXECUTE><<<grid, kThreads, ctx.smem_bytes, e->stream>>>(a);
  else if (mode == MODE_QUERY)
    cw_batch_kernel<MODE_QUERY><<<grid, kThreads, ctx.smem_bytes, e->stream>>>(a);
  else if (mode == MODE_RULES)
    cw_batch_kernel<MODE_RULES><<<grid, kThreads, ctx.smem_bytes, e->stream>>>(a);
  else
    cw_batch_kernel<MODE_REPLAY><<<grid, kThreads, ctx.smem_bytes, e->stream>>>(a);
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

}  // namespace brm
