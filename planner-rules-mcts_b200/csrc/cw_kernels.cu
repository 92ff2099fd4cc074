// Continuous-world kernels (sm_100a).  One kernel template, cw_pool_kernel<MODE>:
//   ROLLOUT  K3  fused random rollouts (mvmcts::RandomHeuristic over MvmctsState::Execute)
//   EXECUTE  K4  one Execute for n states          RULES    MvmctsState::EvaluateRules()
//   REPLAY   K6  given action sequences with a full trace
//   QUERY        CheckTerminal / GetNumActions / GetTerminalReward
//
// A CTA keeps a POOL of P rollouts (states) in shared memory and advances all of them one
// step per iteration.  The step is cut into per-agent work items (cw_core.cuh) that are
// compacted into lists, so that the lanes of a warp always work on agents that have work:
//   S1  one thread per rollout slot: bookkeeping of the last step, finished rollouts are
//       recorded for finalisation and their slot takes the next rollout at once (claimed from
//       a global counter); agents of continuing rollouts are pushed on the phase lists
//   S2  recycle items (slot, agent): fold the finished rollout into the per-leaf sums, load
//       the new leaf's agent (pose, lane projection, automaton states)
//   P1  phase-1 items (slot, agent that moves): action, kinematics, drivable area, projection
//   P2  phase-2 items (slot, MCTS agent): collisions, labels, automata, rewards, terminal flag
// Collided (frozen) agents, finished rollouts and agents that are only simulated cost no lanes:
// every list is dense.  The scenario blob (parameters, lane polylines, road polygon, automaton
// tables) is staged into shared memory by one TMA bulk copy per CTA (re-staged only when the
// CTA moves to a leaf of another scenario).
#include <algorithm>

#include "common.cuh"
#include "cw_core.cuh"
#include "cw_launch.cuh"
#include "tma.cuh"

namespace brm {

namespace {

constexpr int kThreads = kCwThreads;
constexpr double kFixScale = 16777216.0;  // 2^24: fixed-point scale of the per-leaf return sums

// (re)stage the blob of scenario `scn`; all threads of the CTA call this together
__device__ __forceinline__ void cw_stage(uint8_t* smem, uint64_t* bar, const CwArgs& c, int scn, int& cur_scn,
                                         uint32_t& phase) {
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

// append `val` to a shared-memory list for the lanes whose `pred` holds (one atomic per warp);
// must be reached by all 32 lanes of the warp
__device__ __forceinline__ void cw_push(uint16_t* list, int* cnt, bool pred, uint16_t val) {
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  if (!mask) return;
  const int lane = threadIdx.x & 31, leader = __ffs(static_cast<int>(mask)) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(cnt, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) list[base + __popc(mask & ((1u << lane) - 1u))] = val;
}

// which agents of a live rollout go on the phase lists, by mode
template <int MODE>
__device__ __forceinline__ bool cw_in_list1(uint8_t f) {
  if (MODE == MODE_RULES || MODE == MODE_QUERY) return false;
  return (f & BRM_CW_PRESENT) && (f & BRM_CW_VALID);
}
template <int MODE>
__device__ __forceinline__ bool cw_in_list2(uint8_t f, int j) {
  if (MODE == MODE_QUERY) return j == 0;
  if (MODE == MODE_RULES) return (f & BRM_CW_PRESENT) && (f & BRM_CW_VALID);
  return (f & BRM_CW_PRESENT) && ((f & BRM_CW_VALID) || j == 0);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, kCwCtasPerSm) cw_pool_kernel(const CwArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_cnt[2][4];  // per iteration parity: items on list 1, list 2, recycle list
  __shared__ int s_any, s_next, s_end;
  __shared__ unsigned long long s_unit;

  const int tid = threadIdx.x;
  const int P = args.P, A = args.A, M = args.M, R = args.R, V = args.V;
  const CwBlob& B = *reinterpret_cast<const CwBlob*>(smem);
  CwPool p;
  cw_pool_carve(p, smem + args.blob_cap, P, A, M, R, V);
  // scheduler arrays behind the pool
  uint8_t* sp = smem + args.blob_cap + cw_pool_bytes(P, A, M, R);
  double* fin_disc = reinterpret_cast<double*>(sp); sp += static_cast<size_t>(P) * 8;
  int32_t* fin_flat = reinterpret_cast<int32_t*>(sp); sp += static_cast<size_t>(P) * 4;
  int32_t* fin_leaf = reinterpret_cast<int32_t*>(sp); sp += static_cast<size_t>(P) * 4;
  int32_t* fin_k = reinterpret_cast<int32_t*>(sp); sp += static_cast<size_t>(P) * 4;
  int32_t* fin_horizon = reinterpret_cast<int32_t*>(sp); sp += static_cast<size_t>(P) * 4;
  uint32_t* fin_nsteps = reinterpret_cast<uint32_t*>(sp); sp += static_cast<size_t>(P) * 4;
  uint16_t* list1 = reinterpret_cast<uint16_t*>(sp); sp += static_cast<size_t>(P) * A * 2;
  uint16_t* list2 = reinterpret_cast<uint16_t*>(sp); sp += static_cast<size_t>(P) * M * 2;
  uint16_t* listR = reinterpret_cast<uint16_t*>(sp); sp += static_cast<size_t>(P) * A * 2;
  uint8_t* rflag = sp; sp += P;
  uint8_t* livef = sp; sp += P;
  uint8_t* fin_done = sp; sp += P;

  const bool single = args.single_scenario != 0;
  const int n_per_leaf = args.rollouts_per_leaf;
  const int max_steps = args.max_steps;
  const double gamma = args.gamma;
  const double disc0 = (MODE == MODE_ROLLOUT && args.first_exp) ? gamma : 1.0;
  constexpr bool kSteps = MODE == MODE_ROLLOUT || MODE == MODE_EXECUTE || MODE == MODE_REPLAY;

  CwStatesView view;
  view.agents = args.in.agents;
  view.flags = args.in.flags;
  view.q = args.in.q;
  view.n = args.in.n;
  CwStepCtx ctx;
  ctx.seed = args.seed;
  ctx.rollout_base = args.rollout_base;
  ctx.actions = MODE == MODE_ROLLOUT ? nullptr : args.actions;
  ctx.n = args.in.n;
  ctx.T = MODE == MODE_REPLAY ? args.T : 0;

  if (tid < P) {
    p.busy[tid] = 0;
    livef[tid] = 0;
  }
  if (tid == 0) {
    s_cnt[0][0] = s_cnt[0][1] = s_cnt[0][2] = 0;
    s_cnt[1][0] = s_cnt[1][1] = s_cnt[1][2] = 0;
    s_any = 0;
    s_next = 0;
    s_end = 0;
  }
  int cur_scn = -1;
  uint32_t tma_phase = 0;
  int unit_leaf = 0;
  bool need_unit = !single;
  if (single) cw_stage(smem, &bar, args, MODE == MODE_ROLLOUT ? 0 : args.in.scenario[0], cur_scn, tma_phase);
  __syncthreads();

  // next unclaimed work unit for the lanes that want one (-1: none left)
  auto claim = [&](bool want) -> long long {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return -1;
    const int lane = tid & 31, leader = __ffs(static_cast<int>(mask)) - 1;
    long long base = 0;
    if (lane == leader) {
      if (single)
        base = static_cast<long long>(atomicAdd(args.work_counter, static_cast<unsigned long long>(__popc(mask))));
      else
        base = atomicAdd(&s_next, __popc(mask));
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!want) return -1;
    const long long id = base + __popc(mask & ((1u << lane) - 1u));
    if (single) return id < args.total ? id : -1;
    return id < s_end ? static_cast<long long>(unit_leaf) * n_per_leaf + id : -1;
  };

  // ---- what is written when a rollout (state) leaves the pool
  auto finalize = [&](int r, int j) {
    const int s = r * A + j, sm = r * M + j;
    const int64_t flat = fin_flat[r];
    const int64_t leaf = fin_leaf[r];
    if (MODE == MODE_ROLLOUT) {
      if (j < M) {
        double tr[4];
        cw_terminal_reward(B, p, r, j, tr);
#pragma unroll
        for (int v = 0; v < BRM_MAX_REWARD; ++v) {
          if (v >= V) break;
          const double ret = fma(fin_disc[r], tr[v], p.acc[sm * 4 + v]);
          // fixed point per rollout: the integer sums do not depend on the order of the atomics
          const long long fx = __double2ll_rn(ret * kFixScale);
          if (fx)
            atomicAdd(reinterpret_cast<unsigned long long*>(&args.returns_fx[(leaf * M + j) * V + v]),
                      static_cast<unsigned long long>(fx));
          if (args.out.ro_returns) args.out.ro_returns[(flat * M + j) * V + v] = ret;
        }
        for (int sl = 0; sl < R; ++sl) {
          const uint8_t vv = p.viol[sm * R + sl];
          if (vv && args.out.viol_sum)
            atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.viol_sum[(leaf * M + j) * R + sl]),
                      static_cast<unsigned long long>(vv));
          if (args.out.ro_viol) args.out.ro_viol[(flat * M + j) * R + sl] = vv;
          if (args.out.ro_q) args.out.ro_q[(flat * M + j) * R + sl] = static_cast<uint8_t>(cw_q_get(p.qw + sm * p.QW, sl));
        }
      }
      if (args.out.ro_agents) {
        double* o = args.out.ro_agents + (flat * A + j) * 4;
        reinterpret_cast<double2*>(o)[0] = make_double2(p.x[s], p.y[s]);
        reinterpret_cast<double2*>(o)[1] = make_double2(p.th[s], p.v[s]);
      }
      if (args.out.ro_flags) args.out.ro_flags[flat * A + j] = p.flags[s];
      if (j == 0) {
        if (args.out.agent_steps && fin_nsteps[r])
          atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.agent_steps[leaf]),
                    static_cast<unsigned long long>(fin_nsteps[r]));
        if (args.out.ro_steps) args.out.ro_steps[flat] = fin_k[r];
      }
    } else if (MODE == MODE_EXECUTE || MODE == MODE_RULES) {
      const int64_t n = args.in.n, i = leaf;
      double* o = args.bout.agents + (static_cast<int64_t>(j) * n + i) * 4;
      reinterpret_cast<double2*>(o)[0] = make_double2(p.x[s], p.y[s]);
      reinterpret_cast<double2*>(o)[1] = make_double2(p.th[s], p.v[s]);
      args.bout.flags[static_cast<int64_t>(j) * n + i] = p.flags[s];
      if (j == 0) {
        args.bout.horizon[i] = fin_horizon[r];
        // CheckTerminal of the new state; "no ego" (:267-270) is not seen by any work item
        args.bout.terminal[i] = MODE == MODE_RULES ? args.in.terminal[i]
                                                   : ((fin_done[r] || !(p.flags[s] & BRM_CW_PRESENT)) ? 1 : 0);
        args.bout.scenario[i] = args.in.scenario[i];
      }
      if (j < M) {
        for (int v = 0; v < V; ++v) args.rewards[(static_cast<int64_t>(j) * V + v) * n + i] = p.acc[sm * 4 + v];
        for (int sl = 0; sl < R; ++sl) {
          const int64_t off = (static_cast<int64_t>(j) * R + sl) * n + i;
          args.bout.q[off] = static_cast<uint8_t>(cw_q_get(p.qw + sm * p.QW, sl));
          if (args.bout.viol) args.bout.viol[off] = (args.in.viol ? args.in.viol[off] : 0u) + p.viol[sm * R + sl];
        }
      }
    } else if (MODE == MODE_QUERY) {
      const int64_t n = args.in.n, i = leaf;
      if (j == 0 && args.do_check_terminal) args.in.terminal[i] = fin_done[r];
      if (j < M) {
        if (args.num_actions) {
          const uint8_t f = p.flags[s];
          args.num_actions[static_cast<int64_t>(j) * n + i] =
              ((f & BRM_CW_PRESENT) && (f & BRM_CW_VALID)) ? B.n_prims[j] : 1;
        }
        if (args.term_reward) {
          double tr[4];
          cw_terminal_reward(B, p, r, j, tr);
          for (int v = 0; v < V; ++v) args.term_reward[(static_cast<int64_t>(j) * V + v) * n + i] = tr[v];
        }
      }
    } else {  // MODE_REPLAY
      const int64_t i = leaf;
      if (j == 0) args.tr.steps_out[i] = fin_k[r];
      if (j < M) {
        double tr[4];
        cw_terminal_reward(B, p, r, j, tr);
        for (int v = 0; v < V; ++v) args.tr.term_reward_out[(i * M + j) * V + v] = tr[v];
      }
    }
  };

  for (int it = 0;; ++it) {
    if (need_unit) {
      // several scenarios: the blob is CTA-wide, so the CTA works on the rollouts of one leaf
      // at a time (claimed from the global counter), the pool drains in between
      if (tid == 0) {
        const unsigned long long u = atomicAdd(args.work_counter, 1ull);
        s_unit = u;
        s_next = 0;
        s_end = n_per_leaf;
        // the pool is drained: both parities of the list counters start from zero
        s_cnt[0][0] = s_cnt[0][1] = s_cnt[0][2] = 0;
        s_cnt[1][0] = s_cnt[1][1] = s_cnt[1][2] = 0;
      }
      __syncthreads();
      if (s_unit >= static_cast<unsigned long long>(args.in.n)) break;
      unit_leaf = static_cast<int>(s_unit);
      cw_stage(smem, &bar, args, args.in.scenario[unit_leaf], cur_scn, tma_phase);
      need_unit = false;
    }
    int* cnt = s_cnt[it & 1];
    // ------------------------------------------------------------ S1: rollout slots
    {
      const int r = tid;
      const bool in_range = r < P;
      bool fin = false, init = false, busy = false, cont_live = false;
      if (in_range) {
        busy = p.busy[r] != 0;
        if (busy) {
          if (livef[r]) {
            p.k[r] += 1;
            if (kSteps) {
              p.disc[r] *= gamma;
              p.horizon[r] -= 1;
            }
          }
          if (p.done[r] || p.k[r] >= max_steps) {
            fin = true;
            fin_flat[r] = p.flat[r];
            fin_leaf[r] = p.leaf[r];
            fin_k[r] = p.k[r];
            fin_horizon[r] = p.horizon[r];
            fin_nsteps[r] = p.nsteps[r];
            fin_disc[r] = p.disc[r];
            fin_done[r] = p.done[r];
          }
        }
      }
      const bool want = in_range && (!busy || fin);
      const long long id = claim(want);
      if (in_range) {
        if (want) {
          if (id >= 0) {
            const int leaf = static_cast<int>(id / n_per_leaf);
            p.flat[r] = static_cast<int>(id);
            p.leaf[r] = leaf;
            p.k[r] = 0;
            p.horizon[r] = args.in.horizon[leaf];
            p.disc[r] = disc0;
            p.nsteps[r] = 0;
            p.busy[r] = 1;
            p.done[r] = (MODE == MODE_ROLLOUT || MODE == MODE_REPLAY) ? (args.in.terminal[leaf] != 0 ? 1 : 0) : 0;
            init = true;
            busy = true;
          } else {
            p.busy[r] = 0;
            busy = false;
          }
        }
        const bool live = busy && !p.done[r] && p.k[r] < max_steps;
        livef[r] = live ? 1 : 0;
        rflag[r] = static_cast<uint8_t>((fin ? 1 : 0) | (init ? 2 : 0));
        cont_live = live && !init;
        if (busy || fin) s_any = 1;
      }
      const uint16_t e0 = static_cast<uint16_t>(r << 4);
      for (int j = 0; j < A; ++j) {
        cw_push(listR, &cnt[2], in_range && (fin || init), static_cast<uint16_t>(e0 | j));
        const uint8_t f = cont_live ? p.flags[r * A + j] : 0;
        if (MODE != MODE_RULES && MODE != MODE_QUERY)
          cw_push(list1, &cnt[0], cont_live && cw_in_list1<MODE>(f), static_cast<uint16_t>(e0 | j));
        if (j < M) cw_push(list2, &cnt[1], cont_live && cw_in_list2<MODE>(f, j), static_cast<uint16_t>(e0 | j));
      }
    }
    __syncthreads();
    if (!s_any) {
      if (single) break;
      need_unit = true;
      __syncthreads();  // everybody has read s_any before thread 0 touches the counters again
      continue;
    }
    // ------------------------------------------------------------ S2: recycle items
    {
      const int nR = cnt[2];
      for (int i0 = 0; i0 < nR; i0 += kThreads) {
        const int i = i0 + tid;
        bool p1 = false, p2 = false;
        uint16_t e = 0;
        if (i < nR) {
          e = listR[i];
          const int r = e >> 4, j = e & 15;
          const uint8_t rf = rflag[r];
          if (rf & 1) finalize(r, j);
          if (rf & 2) {
            cw_init_item(B, p, view, r, j, p.leaf[r]);
            if (livef[r]) {
              const uint8_t f = p.flags[r * A + j];
              p1 = cw_in_list1<MODE>(f);
              p2 = j < M && cw_in_list2<MODE>(f, j);
            }
          }
        }
        if (MODE != MODE_RULES && MODE != MODE_QUERY) cw_push(list1, &cnt[0], p1, e);
        cw_push(list2, &cnt[1], p2, e);
      }
    }
    __syncthreads();
    // ------------------------------------------------------------ phase 1
    const int n1 = cnt[0], n2 = cnt[1];
    if (tid == 0) {
      int* nxt = s_cnt[(it & 1) ^ 1];
      nxt[0] = nxt[1] = nxt[2] = 0;
      s_any = 0;
    }
    if (MODE == MODE_RULES) {
      for (int i = tid; i < n2; i += kThreads) {
        const uint16_t e = list2[i];
        cw_shape_bits_item(B, p, e >> 4, e & 15);
      }
    } else if (MODE != MODE_QUERY) {
      for (int i = tid; i < n1; i += kThreads) {
        const uint16_t e = list1[i];
        const int r = e >> 4;
        cw_phase1_item(B, p, ctx, r, e & 15);
        atomicAdd(&p.nsteps[r], 1u);
      }
    }
    __syncthreads();
    // ------------------------------------------------------------ phase 2
    for (int i = tid; i < n2; i += kThreads) {
      const uint16_t e = list2[i];
      const int r = e >> 4, a = e & 15;
      if (MODE == MODE_QUERY) {
        if (args.do_check_terminal) p.done[r] = cw_check_terminal(B, p, r) ? 1 : 0;
      } else {
        CwEval ev;
        cw_phase2_item(B, p, r, a, MODE == MODE_RULES, ev);
        if (MODE == MODE_REPLAY) {
          const int64_t bt = static_cast<int64_t>(p.leaf[r]) * args.T + p.k[r];
          for (int v = 0; v < V; ++v) args.tr.rewards_out[(bt * M + a) * V + v] = ev.r[v];
          for (int w = 0; w < 4; ++w) args.tr.labels_out[(bt * M + a) * 4 + w] = ev.w[w];
        }
      }
    }
    __syncthreads();
    if (MODE == MODE_REPLAY) {
      // the trace of this step: every agent of every rollout that executed it
      for (int s = tid; s < P * A; s += kThreads) {
        const int r = s / A, j = s - r * A;
        if (!livef[r]) continue;
        const int k = p.k[r];
        const int64_t bt = static_cast<int64_t>(p.leaf[r]) * args.T + k;
        double* o = args.tr.agents_out + (bt * A + j) * 4;
        o[0] = p.x[s]; o[1] = p.y[s]; o[2] = p.th[s]; o[3] = p.v[s];
        args.tr.flags_out[bt * A + j] = p.flags[s];
        double* fo = args.tr.frenet_out + (bt * A + j) * 3;
        fo[0] = static_cast<double>(p.lane[s]); fo[1] = p.s[s]; fo[2] = p.lat[s];
        if (j == 0) args.tr.terminal_out[bt] = p.done[r];
        if (j < M) {
          const int sm = r * M + j;
          for (int sl = 0; sl < R; ++sl) {
            const uint32_t prev = k > 0 ? args.tr.viol_out[((bt - 1) * M + j) * R + sl] : 0u;
            args.tr.viol_out[(bt * M + j) * R + sl] = prev + p.viol[sm * R + sl];
            p.viol[sm * R + sl] = 0;
            args.tr.q_out[(bt * M + j) * R + sl] = static_cast<uint8_t>(cw_q_get(p.qw + sm * p.QW, sl));
          }
        }
      }
      __syncthreads();
    }
  }
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

template <int MODE>
int launch_mode(brm_engine* e, const CwLaunchCtx& ctx, CwArgs& a, int64_t units) {
  // the opt-in shared-memory size is a per-device attribute of the function: set it on every
  // launch (an engine per GPU in one process is a supported use)
  BRM_CUDA_CHECK(cudaFuncSetAttribute(cw_pool_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(ctx.smem_bytes)));
  int per_sm = 0;
  BRM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cw_pool_kernel<MODE>, kThreads, ctx.smem_bytes));
  BRM_REQUIRE(per_sm >= 1, "continuous-world kernel does not fit on an SM (%zu bytes of shared memory)",
              ctx.smem_bytes);
  int64_t grid = static_cast<int64_t>(e->sm_count) * per_sm;
  const int64_t need = a.single_scenario ? (units + ctx.P - 1) / ctx.P : units;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  cw_pool_kernel<MODE><<<static_cast<int>(grid), kThreads, ctx.smem_bytes, e->stream>>>(a);
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

void fill_common(const CwLaunchCtx& ctx, CwArgs& a) {
  a.blobs = ctx.blobs;
  a.blob_bytes = ctx.blob_bytes;
  a.blob_stride = ctx.blob_stride;
  a.blob_cap = ctx.blob_cap;
  a.A = ctx.A; a.M = ctx.M; a.V = ctx.V; a.R = ctx.R; a.P = ctx.P;
}

}  // namespace

int cw_launch_rollout(brm_engine* e, const CwLaunchCtx& ctx, int n_scenarios, const brm_cw_states& leaves,
                      const brm_rollout_params& rp, const brm_cw_rollout_out& out) {
  const int M = ctx.M, V = ctx.V, R = ctx.R;
  CwArgs a;
  memset(&a, 0, sizeof(a));
  fill_common(ctx, a);
  a.in = leaves;
  a.out = out;
  a.seed = rp.seed;
  a.rollout_base = rp.rollout_base;
  a.rollouts_per_leaf = rp.rollouts_per_leaf;
  a.max_steps = rp.max_steps;
  a.first_exp = rp.first_discount_exp;
  a.gamma = rp.discount_factor;
  a.single_scenario = n_scenarios == 1 ? 1 : 0;
  a.total = static_cast<int64_t>(leaves.n) * rp.rollouts_per_leaf;
  BRM_REQUIRE(a.total + 4 * 32 < (1ll << 31), "brm_cw_rollout: %lld rollouts in one call (limit 2^31)",
              static_cast<long long>(a.total));
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
    BRM_CUDA_CHECK(cudaMemsetAsync(out.agent_steps, 0, static_cast<size_t>(leaves.n) * sizeof(uint64_t), e->stream));
  e->time_begin();
  rc = launch_mode<MODE_ROLLOUT>(e, ctx, a, a.single_scenario ? a.total : leaves.n);
  if (rc != BRM_OK) return rc;
  cw_finalize_returns_kernel<<<static_cast<int>((fx_count + 255) / 256), 256, 0, e->stream>>>(
      a.returns_fx, out.returns_sum, static_cast<int64_t>(fx_count), M, V, rp.coop_factor);
  e->launches++;
  e->time_end();
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

int cw_launch_batch(brm_engine* e, const CwLaunchCtx& ctx, int mode, const brm_cw_states& in,
                    const brm_cw_states* out, const uint8_t* actions, double* rewards, uint8_t* num_actions,
                    double* term_reward, int do_check_terminal, int32_t T, const brm_cw_trace* tr) {
  CwArgs a;
  memset(&a, 0, sizeof(a));
  fill_common(ctx, a);
  a.in = in;
  if (out) a.bout = *out;
  a.actions = actions;
  a.rewards = rewards;
  a.num_actions = num_actions;
  a.term_reward = term_reward;
  a.do_check_terminal = do_check_terminal;
  a.T = T;
  if (tr) a.tr = *tr;
  a.single_scenario = 1;  // batches are homogeneous in scenario (checked by the host API)
  a.rollouts_per_leaf = 1;
  a.total = in.n;
  a.gamma = 1.0;
  a.max_steps = mode == MODE_REPLAY ? T : 1;
  int rc = e->ensure_ws(sizeof(long long));
  if (rc != BRM_OK) return rc;
  a.work_counter = static_cast<unsigned long long*>(e->d_ws);
  BRM_CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), e->stream));
  if (mode == MODE_EXECUTE) return launch_mode<MODE_EXECUTE>(e, ctx, a, in.n);
  if (mode == MODE_RULES) return launch_mode<MODE_RULES>(e, ctx, a, in.n);
  if (mode == MODE_QUERY) return launch_mode<MODE_QUERY>(e, ctx, a, in.n);
  return launch_mode<MODE_REPLAY>(e, ctx, a, in.n);
}

}  // namespace brm
