// GridWorld kernels (sm_100a):
//   K1 gw_rollout_kernel   fused random rollouts (mvmcts::RandomHeuristic over
//                          GridWorldState::Execute), one thread = one rollout, all
//                          agents in registers, automaton tables in shared memory.
//   K2 gw_execute_kernel   one Execute for n states, SoA in / SoA out.
//   K6 gw_replay_kernel    replay of given action sequences with a full trace.
// The scenario blob (parameters, rule-slot layout, transition tables) is staged into
// shared memory with one TMA bulk copy per CTA.
#include "common.cuh"
#include "gw_device.cuh"
#include "philox.cuh"
#include "tma.cuh"

namespace brm {

namespace {

#ifndef BRM_GW_THREADS
#define BRM_GW_THREADS 256
#endif
#ifndef BRM_GW_MIN_CTAS
#define BRM_GW_MIN_CTAS 2
#endif
constexpr int kRolloutThreads = BRM_GW_THREADS;

template <int A>
__device__ __forceinline__ void gw_load_state(const brm_gw_states& S, int R, int64_t i,
                                              GwRegs<A>& s) {
  const int4* ag = reinterpret_cast<const int4*>(S.agents);
  const int64_t n = S.n;
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const int4 v = ag[a * n + i];
    s.x[a] = v.x;
    s.last[a] = v.y;
    s.lane[a] = v.z;
    s.init_lane[a] = v.w;
    uint64_t q = 0;
    for (int r = 0; r < R; ++r)
      q |= static_cast<uint64_t>(S.q[(static_cast<int64_t>(a) * R + r) * n + i] & 15u) << (4 * r);
    s.q[a] = q;
  }
  s.term = S.term[i];
  s.depth = S.depth[i];
}

// ------------------------------------------------------------------ K1 rollout
struct GwRolloutArgs {
  brm_gw_states leaves;
  brm_gw_rollout_out out;   // final (device) outputs; returns_sum produced by the reduce kernel
  double* part_returns;     // [units][A*V] warp partial sums
  const GwBlob* blob;
  uint64_t seed, rollout_base;
  int32_t rollouts_per_leaf, max_steps, first_exp;
  int32_t warps_per_leaf;
  int64_t units;            // leaves * warps_per_leaf
  double gamma;
};

template <int A>
__global__ void __launch_bounds__(kRolloutThreads, BRM_GW_MIN_CTAS)
gw_rollout_kernel(const GwRolloutArgs args) {
  __shared__ GwBlob sB;
  __shared__ __align__(8) uint64_t bar;
  stage_blob(&sB, args.blob, static_cast<uint32_t>(sizeof(GwBlob)), &bar, 0, true);
  const GwBlob& B = sB;
  const int V = B.V, R = B.R;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int n_act = B.n_actions;

  for (int64_t unit = warp0; unit < args.units; unit += n_warps) {
    const int64_t leaf = unit / args.warps_per_leaf;
    const int chunk = static_cast<int>(unit - leaf * args.warps_per_leaf);
    const int idx = chunk * 32 + lane;  // rollout index within the leaf
    const bool active = idx < args.rollouts_per_leaf;
    const uint64_t local = static_cast<uint64_t>(leaf) * args.rollouts_per_leaf + idx;
    const uint64_t rid = args.rollout_base + local;

    GwRegs<A> s;
    gw_load_state<A>(args.leaves, R, leaf, s);
    double acc[A][BRM_MAX_REWARD];
    uint64_t viol[A][2];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      viol[a][0] = viol[a][1] = 0;
#pragma unroll
      for (int v = 0; v < BRM_MAX_REWARD; ++v) acc[a][v] = 0.0;
    }
    double disc = args.first_exp ? args.gamma : 1.0;
    uint32_t agent_steps = 0;
    int k = 0;
    if (active) {
      Philox4 blk[A];
      while (!(s.term & 1u) && k < args.max_steps) {
        uint32_t act[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
          if ((k & 3) == 0) blk[a] = action_block(args.seed, rid, k, a);
          const bool t = (s.term >> a) & 1u;
          act[a] = draw_action(pick_word(blk[a], k), t ? 1u : static_cast<uint32_t>(n_act));
          agent_steps += t ? 0u : 1u;
        }
        gw_step<A>(
            B, s, act,
            [&](int i, const double(&r)[BRM_MAX_REWARD]) {
#pragma unroll
              for (int v = 0; v < BRM_MAX_REWARD; ++v)
                acc[i][v] = __dadd_rn(acc[i][v], __dmul_rn(disc, r[v]));
            },
            [&](int i, int sl) { viol[i][sl >> 3] += 1ull << (8 * (sl & 7)); },
            [](int, const GwLabels&) {});
        disc = __dmul_rn(disc, args.gamma);
        ++k;
      }
#pragma unroll
      for (int a = 0; a < A; ++a) {
        double tr[BRM_MAX_REWARD];
        gw_terminal_reward<A>(B, s, a, tr);
#pragma unroll
        for (int v = 0; v < BRM_MAX_REWARD; ++v)
          acc[a][v] = __dadd_rn(acc[a][v], __dmul_rn(disc, tr[v]));
      }
      // optional per-rollout outputs (parity instrument)
      if (args.out.ro_returns) {
        for (int a = 0; a < A; ++a)
          for (int v = 0; v < V; ++v) args.out.ro_returns[(local * A + a) * V + v] = acc[a][v];
      }
      if (args.out.ro_viol || args.out.ro_q) {
        for (int a = 0; a < A; ++a)
          for (int r = 0; r < R; ++r) {
            if (args.out.ro_viol)
              args.out.ro_viol[(local * A + a) * R + r] =
                  static_cast<uint8_t>(viol[a][r >> 3] >> (8 * (r & 7)));
            if (args.out.ro_q)
              args.out.ro_q[(local * A + a) * R + r] = static_cast<uint8_t>((s.q[a] >> (4 * r)) & 15u);
          }
      }
      if (args.out.ro_agents) {
        int4* o = reinterpret_cast<int4*>(args.out.ro_agents);
#pragma unroll
        for (int a = 0; a < A; ++a)
          o[local * A + a] = make_int4(s.x[a], s.last[a], s.lane[a], s.init_lane[a]);
      }
      if (args.out.ro_steps) args.out.ro_steps[local] = k;
    }
    // warp-level reduction of the unit (inactive lanes contribute zeros)
#pragma unroll
    for (int a = 0; a < A; ++a) {
#pragma unroll
      for (int v = 0; v < BRM_MAX_REWARD; ++v) {
        if (v < V) {
          double x = acc[a][v];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
          if (lane == 0) args.part_returns[unit * (A * V) + a * V + v] = x;
        }
      }
    }
    if (args.out.viol_sum) {
      for (int a = 0; a < A; ++a)
        for (int r = 0; r < R; ++r) {
          const uint32_t c = static_cast<uint32_t>(viol[a][r >> 3] >> (8 * (r & 7))) & 255u;
          const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
          if (lane == 0 && tot)
            atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.viol_sum[(leaf * A + a) * R + r]),
                      static_cast<unsigned long long>(tot));
        }
    }
    if (args.out.agent_steps) {
      const uint32_t tot = __reduce_add_sync(0xffffffffu, agent_steps);
      if (lane == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.agent_steps[leaf]),
                  static_cast<unsigned long long>(tot));
    }
  }
}

// Deterministic second stage: per leaf, sum the warp partials in a fixed order, then mix the
// agents' sums: value_i = own_i + coop * sum_{j != i} own_j (mvmcts COOP_FACTOR, src/util.cpp:28-29;
// 0 in every test of the reference, so the mixing rule itself is unpinned).
__global__ void __launch_bounds__(128)
leaf_reduce_kernel(const double* __restrict__ part, int32_t warps_per_leaf, int32_t n_agents, int32_t V,
                   double coop, double* __restrict__ out) {
  __shared__ double sm[128];
  __shared__ double own[BRM_GW_MAX_AGENTS * BRM_MAX_REWARD];
  const int64_t leaf = blockIdx.x;
  const int nvals = n_agents * V;
  for (int val = 0; val < nvals; ++val) {
    double x = 0.0;
    for (int w = threadIdx.x; w < warps_per_leaf; w += blockDim.x)
      x += part[(leaf * warps_per_leaf + w) * nvals + val];
    sm[threadIdx.x] = x;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) own[val] = sm[0];
    __syncthreads();
  }
  if (threadIdx.x < nvals) {
    const int v = threadIdx.x % V;
    double others = 0.0;
    for (int j = 0; j < n_agents; ++j)
      if (j * V + v != static_cast<int>(threadIdx.x)) others += own[j * V + v];
    out[leaf * nvals + threadIdx.x] = coop != 0.0 ? own[threadIdx.x] + coop * others : own[threadIdx.x];
  }
}

// ------------------------------------------------------------------ K2 execute
struct GwExecArgs {
  brm_gw_states in, out;
  const uint8_t* actions;  // [A][n]
  double* rewards;         // [A][V][n]
  const GwBlob* blob;
};

#ifndef BRM_GW_EXEC_MIN_CTAS
#define BRM_GW_EXEC_MIN_CTAS 4  /* measured on B200: 4 (64 regs) > 3 > 5 > 6 > 8 */
#endif
template <int A>
__global__ void __launch_bounds__(256, BRM_GW_EXEC_MIN_CTAS) gw_execute_kernel(const GwExecArgs args) {
  __shared__ GwBlob sB;
  __shared__ __align__(8) uint64_t bar;
  stage_blob(&sB, args.blob, static_cast<uint32_t>(sizeof(GwBlob)), &bar, 0, true);
  const GwBlob& B = sB;
  const int V = B.V, R = B.R;
  const int64_t n = args.in.n;
  // This kernel streams every state once (HBM-bound): all loads of a state are issued before
  // anything is stored, through non-aliasing read-only pointers, so that they overlap.
  // in and out must not overlap (include/brm.h).
  const int4* __restrict__ in_ag = reinterpret_cast<const int4*>(args.in.agents);
  const uint8_t* __restrict__ in_q = args.in.q;
  const uint8_t* __restrict__ in_term = args.in.term;
  const int32_t* __restrict__ in_depth = args.in.depth;
  const uint32_t* __restrict__ in_viol = args.in.viol;
  const uint8_t* __restrict__ in_act = args.actions;
  int4* __restrict__ out_ag = reinterpret_cast<int4*>(args.out.agents);
  uint8_t* __restrict__ out_q = args.out.q;
  uint8_t* __restrict__ out_term = args.out.term;
  int32_t* __restrict__ out_depth = args.out.depth;
  uint32_t* __restrict__ out_viol = args.out.viol;
  double* __restrict__ out_rew = args.rewards;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    GwRegs<A> s;
    uint32_t act[A];
    uint32_t inc[A];  // slots of agent a that recorded a violation in this step (one bit each)
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const int4 v = __ldg(&in_ag[a * n + i]);
      s.x[a] = v.x;
      s.last[a] = v.y;
      s.lane[a] = v.z;
      s.init_lane[a] = v.w;
      act[a] = __ldg(&in_act[static_cast<int64_t>(a) * n + i]);
      inc[a] = 0u;
    }
    s.term = __ldg(&in_term[i]);
    s.depth = __ldg(&in_depth[i]);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      uint64_t q = 0;
      for (int r = 0; r < R; ++r)
        q |= static_cast<uint64_t>(__ldg(&in_q[(static_cast<int64_t>(a) * R + r) * n + i]) & 15u) << (4 * r);
      s.q[a] = q;
    }
    gw_step<A>(
        B, s, act,
        [&](int a, const double(&r)[BRM_MAX_REWARD]) {
          for (int v = 0; v < V; ++v) out_rew[(static_cast<int64_t>(a) * V + v) * n + i] = r[v];
        },
        [&](int a, int sl) {
#pragma unroll
          for (int b = 0; b < A; ++b)
            if (b == a) inc[b] |= 1u << sl;
        },
        [](int, const GwLabels&) {});
#pragma unroll
    for (int a = 0; a < A; ++a) {
      out_ag[a * n + i] = make_int4(s.x[a], s.last[a], s.lane[a], s.init_lane[a]);
      for (int r = 0; r < R; ++r)
        out_q[(static_cast<int64_t>(a) * R + r) * n + i] = static_cast<uint8_t>((s.q[a] >> (4 * r)) & 15u);
    }
    out_term[i] = static_cast<uint8_t>(s.term);
    out_depth[i] = s.depth;
    // violation counters (if tracked): carried through plus this step's violations
    if (out_viol) {
#pragma unroll
      for (int a = 0; a < A; ++a)
        for (int r = 0; r < R; ++r) {
          const int64_t o = (static_cast<int64_t>(a) * R + r) * n + i;
          out_viol[o] = (in_viol ? __ldg(&in_viol[o]) : 0u) + ((inc[a] >> r) & 1u);
        }
    }
  }
}

// GetNumActions / IsTerminal / GetTerminalReward for n states
struct GwQueryArgs {
  brm_gw_states s;
  uint8_t* num_actions;  // [A][n] or null
  uint8_t* is_terminal;  // [n] or null
  double* term_reward;   // [A][V][n] or null
  const GwBlob* blob;
};

template <int A>
__global__ void __launch_bounds__(256) gw_query_kernel(const GwQueryArgs args) {
  __shared__ GwBlob sB;
  __shared__ __align__(8) uint64_t bar;
  stage_blob(&sB, args.blob, static_cast<uint32_t>(sizeof(GwBlob)), &bar, 0, true);
  const GwBlob& B = sB;
  const int64_t n = args.s.n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    GwRegs<A> s;
    gw_load_state<A>(args.s, B.R, i, s);
    if (args.is_terminal) args.is_terminal[i] = s.term & 1u;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (args.num_actions)
        args.num_actions[static_cast<int64_t>(a) * n + i] =
            ((s.term >> a) & 1u) ? 1 : static_cast<uint8_t>(B.n_actions);
      if (args.term_reward) {
        double r[BRM_MAX_REWARD];
        gw_terminal_reward<A>(B, s, a, r);
        for (int v = 0; v < B.V; ++v)
          args.term_reward[(static_cast<int64_t>(a) * B.V + v) * n + i] = r[v];
      }
    }
  }
}

// ------------------------------------------------------------------ K6 replay
struct GwReplayArgs {
  brm_gw_states s0;
  const uint8_t* actions;  // [B][T][A]
  int32_t T;
  brm_gw_trace tr;
  const GwBlob* blob;
};

template <int A>
__global__ void __launch_bounds__(128) gw_replay_kernel(const GwReplayArgs args) {
  __shared__ GwBlob sB;
  __shared__ __align__(8) uint64_t bar;
  stage_blob(&sB, args.blob, static_cast<uint32_t>(sizeof(GwBlob)), &bar, 0, true);
  const GwBlob& B = sB;
  const int V = B.V, R = B.R, T = args.T;
  const int64_t nb = args.s0.n;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; b < nb;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    GwRegs<A> s;
    gw_load_state<A>(args.s0, R, b, s);
    uint32_t viol[A][BRM_GW_MAX_SLOTS];
    for (int a = 0; a < A; ++a)
      for (int r = 0; r < BRM_GW_MAX_SLOTS; ++r) viol[a][r] = 0;
    int t = 0;
    for (; t < T && !(s.term & 1u); ++t) {
      const int64_t bt = b * T + t;
      uint32_t act[A];
#pragma unroll
      for (int a = 0; a < A; ++a) act[a] = args.actions[bt * A + a];
      gw_step<A>(
          B, s, act,
          [&](int a, const double(&r)[BRM_MAX_REWARD]) {
            for (int v = 0; v < V; ++v) args.tr.rewards_out[(bt * A + a) * V + v] = r[v];
          },
          [&](int a, int sl) { viol[a][sl] += 1u; },
          [&](int a, const GwLabels& lab) {
            uint32_t* o = args.tr.labels_out + (bt * A + a) * 3;
            o[0] = lab.ego;
            o[1] = lab.merged;
            o[2] = lab.front;
          });
      int4* ag = reinterpret_cast<int4*>(args.tr.agents_out);
      for (int a = 0; a < A; ++a) {
        ag[bt * A + a] = make_int4(s.x[a], s.last[a], s.lane[a], s.init_lane[a]);
        for (int r = 0; r < R; ++r) {
          args.tr.q_out[(bt * A + a) * R + r] = static_cast<uint8_t>((s.q[a] >> (4 * r)) & 15u);
          args.tr.viol_out[(bt * A + a) * R + r] = viol[a][r];
        }
      }
      args.tr.term_out[bt] = static_cast<uint8_t>(s.term);
    }
    args.tr.steps_out[b] = t;
    for (int a = 0; a < A; ++a) {
      double r[BRM_MAX_REWARD];
      gw_terminal_reward<A>(B, s, a, r);
      for (int v = 0; v < V; ++v) args.tr.term_reward_out[(b * A + a) * V + v] = r[v];
    }
  }
}

template <template <int> class Launcher, class... Args>
int dispatch_agents(int A, Args&&... args) {
  switch (A) {
    case 1: return Launcher<1>::run(args...);
    case 2: return Launcher<2>::run(args...);
    case 3: return Launcher<3>::run(args...);
    case 4: return Launcher<4>::run(args...);
    case 5: return Launcher<5>::run(args...);
    case 6: return Launcher<6>::run(args...);
    case 7: return Launcher<7>::run(args...);
    case 8: return Launcher<8>::run(args...);
    default: set_error("GridWorld supports 1..8 agents, got %d", A); return BRM_ERR_INVALID;
  }
}

template <int A>
struct RolloutLauncher {
  static int run(brm_engine* e, const GwRolloutArgs& a, int grid) {
    gw_rollout_kernel<A><<<grid, kRolloutThreads, 0, e->stream>>>(a);
    return BRM_OK;
  }
};
template <int A>
struct ExecLauncher {
  static int run(brm_engine* e, const GwExecArgs& a, int grid) {
    gw_execute_kernel<A><<<grid, 256, 0, e->stream>>>(a);
    return BRM_OK;
  }
};
template <int A>
struct QueryLauncher {
  static int run(brm_engine* e, const GwQueryArgs& a, int grid) {
    gw_query_kernel<A><<<grid, 256, 0, e->stream>>>(a);
    return BRM_OK;
  }
};
template <int A>
struct ReplayLauncher {
  static int run(brm_engine* e, const GwReplayArgs& a, int grid) {
    gw_replay_kernel<A><<<grid, 128, 0, e->stream>>>(a);
    return BRM_OK;
  }
};

int grid_for(const brm_engine* e, int64_t threads_needed, int block, int ctas_per_sm) {
  int64_t blocks = (threads_needed + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(e->sm_count) * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

// second-stage reduction of the GridWorld rollout kernel (own[] in leaf_reduce_kernel is sized
// for BRM_GW_MAX_AGENTS * BRM_MAX_REWARD values: GridWorld shapes only)
static void launch_leaf_reduce(brm_engine* e, const double* part, int32_t per_leaf, int32_t n_agents, int32_t V,
                        double coop, double* out, int32_t n_leaves) {
  leaf_reduce_kernel<<<n_leaves, 128, 0, e->stream>>>(part, per_leaf, n_agents, V, coop, out);
  e->launches++;
}

// ---- launch wrappers used by gw_engine.cu (all pointers are device pointers here) ----

int gw_launch_rollout(brm_engine* e, const brm_gw_states& leaves, const brm_rollout_params& rp,
                      const brm_gw_rollout_out& out, bool sums_zeroed) {
  const int A = e->gw_blob.A, V = e->gw_blob.V;
  GwRolloutArgs a;
  a.leaves = leaves;
  a.out = out;
  a.blob = e->d_gw_blob;
  a.seed = rp.seed;
  a.rollout_base = rp.rollout_base;
  a.rollouts_per_leaf = rp.rollouts_per_leaf;
  a.max_steps = rp.max_steps;
  a.first_exp = rp.first_discount_exp;
  a.gamma = rp.discount_factor;
  a.warps_per_leaf = (rp.rollouts_per_leaf + 31) / 32;
  a.units = static_cast<int64_t>(leaves.n) * a.warps_per_leaf;
  const size_t part_bytes = static_cast<size_t>(a.units) * A * V * sizeof(double);
  int rc = e->ensure_ws(part_bytes);
  if (rc != BRM_OK) return rc;
  a.part_returns = static_cast<double*>(e->d_ws);
  if (out.viol_sum && !sums_zeroed)
    BRM_CUDA_CHECK(cudaMemsetAsync(out.viol_sum, 0,
                                   static_cast<size_t>(leaves.n) * A * e->gw_blob.R * sizeof(uint64_t),
                                   e->stream));
  if (out.agent_steps && !sums_zeroed)
    BRM_CUDA_CHECK(cudaMemsetAsync(out.agent_steps, 0, static_cast<size_t>(leaves.n) * sizeof(uint64_t),
                                   e->stream));
  // persistent grid: a multiple of the SM count, 4 CTAs of 8 warps per SM
  const int grid = grid_for(e, a.units * 32, kRolloutThreads, BRM_GW_MIN_CTAS);
  e->time_begin();
  rc = dispatch_agents<RolloutLauncher>(A, e, a, grid);
  if (rc != BRM_OK) return rc;
  e->launches++;
  launch_leaf_reduce(e, a.part_returns, a.warps_per_leaf, A, V, rp.coop_factor, out.returns_sum, leaves.n);
  e->time_end();
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

int gw_launch_execute(brm_engine* e, const brm_gw_states& in, const uint8_t* actions,
                      const brm_gw_states& out, double* rewards) {
  GwExecArgs a;
  a.in = in;
  a.out = out;
  a.actions = actions;
  a.rewards = rewards;
  a.blob = e->d_gw_blob;
  const int grid = grid_for(e, in.n, 256, BRM_GW_EXEC_MIN_CTAS * 4);
  int rc = dispatch_agents<ExecLauncher>(e->gw_blob.A, e, a, grid);
  if (rc != BRM_OK) return rc;
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

int gw_launch_query(brm_engine* e, const brm_gw_states& s, uint8_t* num_actions,
                    uint8_t* is_terminal, double* term_reward) {
  GwQueryArgs a;
  a.s = s;
  a.num_actions = num_actions;
  a.is_terminal = is_terminal;
  a.term_reward = term_reward;
  a.blob = e->d_gw_blob;
  const int grid = grid_for(e, s.n, 256, 8);
  int rc = dispatch_agents<QueryLauncher>(e->gw_blob.A, e, a, grid);
  if (rc != BRM_OK) return rc;
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

int gw_launch_replay(brm_engine* e, const brm_gw_states& s0, const uint8_t* actions, int32_t T,
                     const brm_gw_trace& tr) {
  GwReplayArgs a;
  a.s0 = s0;
  a.actions = actions;
  a.T = T;
  a.tr = tr;
  a.blob = e->d_gw_blob;
  const int grid = grid_for(e, s0.n, 128, 8);
  int rc = dispatch_agents<ReplayLauncher>(e->gw_blob.A, e, a, grid);
  if (rc != BRM_OK) return rc;
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

}  // namespace brm
