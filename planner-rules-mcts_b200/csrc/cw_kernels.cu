// Continuous-world kernels (sm_100a).  One kernel template, cw_pool_kernel<MODE>:
//   ROLLOUT  K3  fused random rollouts (mvmcts::RandomHeuristic over MvmctsState::Execute)
//   EXECUTE  K4  one Execute for n states          RULES    MvmctsState::EvaluateRules()
//   REPLAY   K6  given action sequences with a full trace
//   QUERY        CheckTerminal / GetNumActions / GetTerminalReward
//
// Every WARP keeps a pool of up to 32 rollouts (states) in its own piece of shared memory and
// advances all of them one step per iteration (the warps of the large CTA shape meet at one
// barrier per iteration, see below; data never crosses warps).  The step is
// cut into per-agent work items (cw_core.cuh) that are compacted into lists, so that the
// lanes always work on agents that have work:
//   S1  lane = rollout slot: bookkeeping of the last step; finished rollouts are recorded for
//       finalisation and free slots take new rollouts (claimed from a global counter) -- as
//       many as keep the number of phase-1 items at the warp's target, a multiple of 32;
//       agents of continuing rollouts are pushed on the phase lists
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

// The warps of a CTA own their pools, but they start the phases of an iteration together (ONE
// barrier per iteration, in front of phase 1, per group of BRM_CW_SYNC_GROUP warps): warps
// spread over all phases need the whole loop (~70 KB of code) in the instruction cache at once
// (a quarter of all stall samples were instruction fetches without it).  The warps have about
// the same number of work items per phase (the refill rule), so they stay close until the next
// barrier; more barriers (one per phase: -4 %) or fewer (every second iteration) measured worse;
// two groups of warps with a named barrier each beat one barrier for all (+3.5 %: less waiting
// for the slowest warp) and groups of 4 or 5.  The barrier doubles as the group's vote on whether
// anybody has work left: a warp whose pool has run dry attends until its whole group is done.
#ifndef BRM_CW_SYNC_GROUP
#define BRM_CW_SYNC_GROUP 7
#endif
#ifndef BRM_CW_PHASE_SYNC
#define BRM_CW_PHASE_SYNC 1
#endif

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

// append `val` to a warp-private list for the lanes whose `pred` holds; `n` is the list length,
// kept in a warp-uniform register.  Must be reached by all 32 lanes.
__device__ __forceinline__ void cw_push(uint16_t* list, int& n, bool pred, uint16_t val) {
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  if (pred) list[n + __popc(mask & ((1u << (threadIdx.x & 31)) - 1u))] = val;
  n += __popc(mask);
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

// agent j of `leaf`, projected and packed the way a pool slot holds it, into the warp's cache
__device__ __forceinline__ void cw_cache_fill(const CwBlob& B, const CwStatesView& S, const CwPool& p,
                                              double* cache_ag, uint32_t* cache_meta, uint32_t* cache_mc, int j,
                                              int64_t leaf) {
  const double* src = S.agents + (static_cast<int64_t>(j) * S.n + leaf) * 4;
  CwKin t;
  t.x = src[0]; t.y = src[1]; t.th = src[2]; t.v = src[3];
  const uint8_t fl = S.flags[static_cast<int64_t>(j) * S.n + leaf];
  cw_sincos(t.th, &t.sn, &t.cs);
  CwFrenet f;
  f.lane = -1; f.s = 0.0; f.lat = 0.0;
  if (fl & BRM_CW_PRESENT) f = cw_frenet(B, t.x, t.y);
  uint8_t bits = 0;
  if (j == 0 && (fl & BRM_CW_PRESENT) && !(fl & BRM_CW_VALID)) {
    double cx[4], cy[4];
    cw_corners(B, t, j, cx, cy);
    if (cw_out_of_map(B, t, j, cx, cy)) bits |= CW_BIT_OOM;
    if (cw_at_goal(B, t, j, cx, cy)) bits |= CW_BIT_GOAL;
  }
  bits |= cw_beyond_bit(B, f);
  double* rec = cache_ag + j * kAgWords;
  rec[AG_X] = t.x; rec[AG_Y] = t.y; rec[AG_TH] = t.th; rec[AG_V] = t.v; rec[AG_CS] = t.cs; rec[AG_SN] = t.sn;
  rec[AG_S] = f.s; rec[AG_LAT] = f.lat; rec[AG_COST] = 0.0;
  CwMeta m;
  m.lane = static_cast<int16_t>(f.lane); m.flags = fl; m.bits = bits;
  *reinterpret_cast<CwMeta*>(cache_meta + j) = m;
  if (j < p.M) {
    uint32_t* w = cache_mc + j * (1 + p.QW);  // live, qw[QW]
    for (int k = 0; k < p.QW; ++k) w[1 + k] = 0u;
    uint32_t live = 0u;
    for (int sl = 0; sl < p.R; ++sl) {
      const uint32_t q = S.q[(static_cast<int64_t>(j) * p.R + sl) * S.n + leaf];
      cw_q_set(w + 1, sl, q);
      if (!cw_slot_absorbing(B, sl, q)) live |= 1u << sl;
    }
    w[0] = live;
  }
}

// slot (r, j) from the cached leaf
__device__ __forceinline__ void cw_init_from_cache(CwPool& p, const double* cache_ag, const uint32_t* cache_meta,
                                                   const uint32_t* cache_mc, int r, int j) {
  const double* src = cache_ag + j * kAgWords;
  double* rec = cw_ag(p, cw_slot(p, r, j));
#pragma unroll
  for (int k = 0; k < kAgWords; ++k) rec[k] = src[k];
  p.meta[cw_slot(p, r, j)] = cache_meta[j];
  if (j < p.M) {
    const int sm = cw_mslot(p, r, j);
    double* mc = cw_mc(p, sm);
    for (int k = 0; k < p.V + 2; ++k) mc[k] = 0.0;  // acc, rng
    const uint32_t* w = cache_mc + j * (1 + p.QW);
    uint32_t* dst = cw_mc_live(p, sm);
    for (int k = 0; k < 1 + p.QW; ++k) dst[k] = w[k];
    uint8_t* viol = cw_mc_viol(p, sm);
    for (int sl = 0; sl < p.R; ++sl) viol[sl] = 0;
  }
}

template <int MODE, int NT, int CTAS>
__global__ void __launch_bounds__(NT, CTAS) cw_pool_kernel(const CwArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_leaf;
  __shared__ unsigned long long s_unit;

  // one CTA per SM (the large shape): its warps are the only ones sharing the instruction cache
  constexpr bool kPhaseSync = BRM_CW_PHASE_SYNC != 0 && NT >= 256;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int P = args.P, A = args.A, M = args.M, R = args.R, V = args.V;
  const unsigned full = 0xffffffffu;
  const CwBlob& B = *reinterpret_cast<const CwBlob*>(smem);
  // this warp's piece of shared memory: pool, then what a finished rollout leaves behind for
  // its finalisation, then the three item lists
  uint8_t* wmem = smem + args.blob_cap + static_cast<size_t>(wid) * args.warp_bytes;
  CwPool p;
  cw_pool_carve(p, wmem, P, A, M, R, V);
  uint8_t* sp = wmem + cw_pool_bytes(P, A, M, R, V);
  double* fin_disc = reinterpret_cast<double*>(sp); sp += static_cast<size_t>(P) * 8;
  int32_t* fin_rec = reinterpret_cast<int32_t*>(sp); sp += static_cast<size_t>(P) * 6 * 4;  // flat, leaf, k, horizon, nsteps, done
  // the leaf most rollouts of the warp's current chunk start from, ready to be copied into a slot
  double* cache_ag = reinterpret_cast<double*>(sp); sp += static_cast<size_t>(A) * kAgWords * 8;
  uint32_t* cache_meta = reinterpret_cast<uint32_t*>(sp); sp += static_cast<size_t>(A) * 4;
  uint32_t* cache_mc = reinterpret_cast<uint32_t*>(sp); sp += static_cast<size_t>(M) * (1 + p.QW) * 4;
  uint16_t* list1 = reinterpret_cast<uint16_t*>(sp); sp += static_cast<size_t>(P) * A * 2;
  uint16_t* list2 = reinterpret_cast<uint16_t*>(sp); sp += static_cast<size_t>(P) * M * 2;
  uint16_t* listR = reinterpret_cast<uint16_t*>(sp);  // recycle items; reused for the pending shape tests
  int cache_leaf = -1;                                // warp-uniform
  long long c_lo = 0, c_hi = 0;                       // warp-uniform: claimed, not yet handed out

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

  if (lane < P) {
    uint8_t* hs = cw_hstate(p, lane);
    hs[HS_BUSY] = 0;
    hs[HS_DONE] = 0;
    hs[HS_LIVE] = 0;
    hs[HS_RFLAG] = 0;
  }
  int cur_scn = -1;
  uint32_t tma_phase = 0;
  int unit_leaf = 0;
  if (single) cw_stage(smem, &bar, args, MODE == MODE_ROLLOUT ? 0 : args.in.scenario[0], cur_scn, tma_phase);
  __syncthreads();

  // ---- what is written when a rollout (state) leaves the pool
  auto finalize = [&](int r, int j) {
    const int s = cw_slot(p, r, j), sm = cw_mslot(p, r, j);
    const CwHdrRef fr{fin_rec + r, P};  // field by field, like the headers
    const int64_t flat = fr[0];
    const int64_t leaf = fr[1];
    const double* rec = cw_ag(p, s);
    const uint8_t flags = cw_meta(p, s)->flags;
    if (MODE == MODE_ROLLOUT) {
      if (j < M) {
        double tr[4];
        cw_terminal_reward(B, p, r, j, tr);
        const double* mc = cw_mc(p, sm);
#pragma unroll
        for (int v = 0; v < BRM_MAX_REWARD; ++v) {
          if (v >= V) break;
          const double ret = fma(fin_disc[r], tr[v], mc[MC_ACC + v]);
          // fixed point per rollout: the integer sums do not depend on the order of the atomics
          const long long fx = __double2ll_rn(ret * kFixScale);
          if (fx)
            atomicAdd(reinterpret_cast<unsigned long long*>(&args.returns_fx[(leaf * M + j) * V + v]),
                      static_cast<unsigned long long>(fx));
          if (args.out.ro_returns) args.out.ro_returns[(flat * M + j) * V + v] = ret;
        }
        const uint8_t* viol = cw_mc_viol(p, sm);
        const uint32_t* qw = cw_mc_qw(p, sm);
        for (int sl = 0; sl < R; ++sl) {
          const uint8_t vv = viol[sl];
          if (vv && args.out.viol_sum)
            atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.viol_sum[(leaf * M + j) * R + sl]),
                      static_cast<unsigned long long>(vv));
          if (args.out.ro_viol) args.out.ro_viol[(flat * M + j) * R + sl] = vv;
          if (args.out.ro_q) args.out.ro_q[(flat * M + j) * R + sl] = static_cast<uint8_t>(cw_q_get(qw, sl));
        }
      }
      if (args.out.ro_agents) {
        double* o = args.out.ro_agents + (flat * A + j) * 4;
        reinterpret_cast<double2*>(o)[0] = make_double2(rec[AG_X], rec[AG_Y]);
        reinterpret_cast<double2*>(o)[1] = make_double2(rec[AG_TH], rec[AG_V]);
      }
      if (args.out.ro_flags) args.out.ro_flags[flat * A + j] = flags;
      if (j == 0) {
        if (args.out.agent_steps && fr[4])
          atomicAdd(reinterpret_cast<unsigned long long*>(&args.out.agent_steps[leaf]),
                    static_cast<unsigned long long>(static_cast<uint32_t>(fr[4])));
        if (args.out.ro_steps) args.out.ro_steps[flat] = fr[2];
      }
    } else if (MODE == MODE_EXECUTE || MODE == MODE_RULES) {
      const int64_t n = args.in.n, i = leaf;
      double* o = args.bout.agents + (static_cast<int64_t>(j) * n + i) * 4;
      reinterpret_cast<double2*>(o)[0] = make_double2(rec[AG_X], rec[AG_Y]);
      reinterpret_cast<double2*>(o)[1] = make_double2(rec[AG_TH], rec[AG_V]);
      args.bout.flags[static_cast<int64_t>(j) * n + i] = flags;
      if (j == 0) {
        args.bout.horizon[i] = fr[3];
        // CheckTerminal of the new state; "no ego" (:267-270) is not seen by any work item
        args.bout.terminal[i] = MODE == MODE_RULES ? args.in.terminal[i]
                                                   : ((fr[5] || !(flags & BRM_CW_PRESENT)) ? 1 : 0);
        args.bout.scenario[i] = args.in.scenario[i];
      }
      if (j < M) {
        const double* mc = cw_mc(p, sm);
        const uint8_t* viol = cw_mc_viol(p, sm);
        const uint32_t* qw = cw_mc_qw(p, sm);
        for (int v = 0; v < V; ++v) args.rewards[(static_cast<int64_t>(j) * V + v) * n + i] = mc[MC_ACC + v];
        for (int sl = 0; sl < R; ++sl) {
          const int64_t off = (static_cast<int64_t>(j) * R + sl) * n + i;
          args.bout.q[off] = static_cast<uint8_t>(cw_q_get(qw, sl));
          if (args.bout.viol) args.bout.viol[off] = (args.in.viol ? args.in.viol[off] : 0u) + viol[sl];
        }
      }
    } else if (MODE == MODE_QUERY) {
      const int64_t n = args.in.n, i = leaf;
      if (j == 0 && args.do_check_terminal) args.in.terminal[i] = fr[5] ? 1 : 0;
      if (j < M) {
        if (args.num_actions)
          args.num_actions[static_cast<int64_t>(j) * n + i] =
              ((flags & BRM_CW_PRESENT) && (flags & BRM_CW_VALID)) ? B.n_prims[j] : 1;
        if (args.term_reward) {
          double tr[4];
          cw_terminal_reward(B, p, r, j, tr);
          for (int v = 0; v < V; ++v) args.term_reward[(static_cast<int64_t>(j) * V + v) * n + i] = tr[v];
        }
      }
    } else {  // MODE_REPLAY
      const int64_t i = leaf;
      if (j == 0) args.tr.steps_out[i] = fr[2];
      if (j < M) {
        double tr[4];
        cw_terminal_reward(B, p, r, j, tr);
        for (int v = 0; v < V; ++v) args.tr.term_reward_out[(i * M + j) * V + v] = tr[v];
      }
    }
  };

  for (;;) {
    if (!single) {
      // several scenarios: the blob is CTA-wide, so the CTA's warps work on the rollouts of one
      // leaf at a time; the pools drain in between.  Every leaf is first taken by one CTA (the
      // global counter); CTAs that find no fresh leaf join the leaves that still have unclaimed
      // rollouts (per-leaf counters in global memory), so the last leaves are shared
      __syncthreads();  // every warp has drained its pool
      if (tid == 0) {
        s_unit = atomicAdd(args.work_counter, 1ull);
        s_leaf = 0x7fffffff;
      }
      __syncthreads();
      const int n_leaves = static_cast<int>(args.in.n);
      if (s_unit < static_cast<unsigned long long>(n_leaves)) {
        unit_leaf = static_cast<int>(s_unit);
      } else {
        const int start = static_cast<int>((s_unit * 2654435761ull) % static_cast<unsigned long long>(n_leaves));
        for (int k = tid; k < n_leaves; k += NT) {
          int l = start + k;
          if (l >= n_leaves) l -= n_leaves;
          if (*reinterpret_cast<volatile int32_t*>(args.leaf_next + l) < n_per_leaf) {
            atomicMin(&s_leaf, k);
            break;
          }
        }
        __syncthreads();
        if (s_leaf == 0x7fffffff) break;  // every rollout of every leaf has been claimed
        unit_leaf = start + s_leaf;
        if (unit_leaf >= n_leaves) unit_leaf -= n_leaves;
      }
      cw_stage(smem, &bar, args, args.in.scenario[unit_leaf], cur_scn, tma_phase);
    }
    for (;;) {
      int n1 = 0, n2 = 0, nR = 0;
      bool has_work = false;
      // ---------------------------------------------------------- S1: lane = rollout slot
      {
        const int r = lane;
        const bool in_range = r < P;
        const CwHdrRef h = cw_hdr(p, in_range ? r : 0);
        uint8_t* hs = cw_hstate(p, in_range ? r : 0);
        bool fin = false, busy = false;
        if (in_range) {
          busy = hs[HS_BUSY] != 0;
          if (busy) {
            if (hs[HS_LIVE]) {
              h[H_K] += 1;
              if (kSteps) {
                p.disc[r] *= gamma;
                h[H_HORIZON] -= 1;
              }
            }
            if (hs[HS_DONE] || h[H_K] >= max_steps) {
              fin = true;
              const CwHdrRef fr{fin_rec + r, P};
              fr[0] = h[H_FLAT]; fr[1] = h[H_LEAF]; fr[2] = h[H_K]; fr[3] = h[H_HORIZON]; fr[4] = h[H_NSTEPS];
              fr[5] = hs[HS_DONE];
              fin_disc[r] = p.disc[r];
            }
          }
        }
        // agents of the rollouts that go on: phase lists (and the agent-steps they will add)
        const bool cont = in_range && busy && !fin;
        const uint16_t e0 = static_cast<uint16_t>(r << 4);
        int moved = 0;
        for (int j = 0; j < A; ++j) {
          const uint8_t f = cont ? cw_meta(p, cw_slot(p, r, j))->flags : 0;
          const bool p1 = cont && cw_in_list1<MODE>(f);
          if (MODE != MODE_RULES && MODE != MODE_QUERY) cw_push(list1, n1, p1, static_cast<uint16_t>(e0 | j));
          moved += p1 ? 1 : 0;
          if (j < M) cw_push(list2, n2, cont && cw_in_list2<MODE>(f, j), static_cast<uint16_t>(e0 | j));
        }
        if (cont) {
          h[H_NSTEPS] += moved;
          hs[HS_LIVE] = 1;
          hs[HS_RFLAG] = 0;
        }
        // free slots take new work: as many as keep the phase-1 items at the target
        const bool want = in_range && (!busy || fin);
        const unsigned wmask = __ballot_sync(full, want);
        // how many of the free slots to fill: the phases run in rounds of 32 items, so aim at the
        // number of FULL rounds the pool could sustain if every free slot were filled (A items
        // per new rollout) -- a pool whose agents froze keeps fewer rollouts in flight rather than
        // a last round with a few lanes
        int allow = __popc(wmask);
        if (MODE == MODE_ROLLOUT && allow > 0) {
          int target = ((n1 + allow * A) >> 5) << 5;
          if (target < 32) target = 32;
          const int k = target > n1 ? (target - n1) / A : 0;
          allow = k < allow ? k : allow;
          if (allow < 1 && n1 == 0) allow = 1;
        }
        // ids come from the warp's chunk of consecutive ids (mostly one leaf: its projected pose
        // is cached); an empty chunk is replaced from the global counter, in sizes that shrink
        // towards the end of the work
        int n_take = min(__popc(wmask), allow);
        if (n_take > 0 && c_lo >= c_hi) {
          long long base = 0, size = 0;
          if (lane == 0) {
            if (single) {
              const long long left = args.total - static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(args.work_counter));
              size = left / (2ll * args.n_warps_total);
              size = size < 4 ? 4 : (size > 64 ? 64 : size);
              base = static_cast<long long>(atomicAdd(args.work_counter, static_cast<unsigned long long>(size)));
              size = base + size > args.total ? (args.total > base ? args.total - base : 0) : size;
            } else {
              size = P;
              const int b = atomicAdd(args.leaf_next + unit_leaf, static_cast<int>(size));
              base = static_cast<long long>(unit_leaf) * n_per_leaf + b;
              size = b + size > n_per_leaf ? (n_per_leaf > b ? n_per_leaf - b : 0) : size;
            }
          }
          c_lo = __shfl_sync(full, base, 0);
          c_hi = c_lo + __shfl_sync(full, size, 0);
        }
        if (c_hi - c_lo < n_take) n_take = static_cast<int>(c_hi - c_lo);
        const int my_rank = __popc(wmask & ((1u << lane) - 1u));
        const bool take = want && my_rank < n_take;
        const long long id = take ? c_lo + my_rank : -1;
        c_lo += n_take;
        bool init = false;
        if (want) {
          if (id >= 0) {
            const int leaf = static_cast<int>(id / n_per_leaf);
            h[H_FLAT] = static_cast<int>(id);
            h[H_LEAF] = leaf;
            h[H_K] = 0;
            h[H_HORIZON] = args.in.horizon[leaf];
            h[H_NSTEPS] = 0;
            p.disc[r] = disc0;
            const bool done0 = (MODE == MODE_ROLLOUT || MODE == MODE_REPLAY) && args.in.terminal[leaf] != 0;
            hs[HS_BUSY] = 1;
            hs[HS_DONE] = done0 ? 1 : 0;
            hs[HS_LIVE] = (!done0 && max_steps > 0) ? 1 : 0;
            init = true;
          } else {
            hs[HS_BUSY] = 0;
            hs[HS_LIVE] = 0;
          }
          hs[HS_RFLAG] = static_cast<uint8_t>((fin ? 1 : 0) | (init ? 2 : 0));
        }
        for (int j = 0; j < A; ++j) cw_push(listR, nR, want && (fin || init), static_cast<uint16_t>(e0 | j));
        has_work = in_range && (busy || init);
        if (!kPhaseSync && !__ballot_sync(full, has_work)) break;  // pool empty, no work left
        // the leaf the new rollouts start from: project it once (lanes = agents), copy it per slot
        const unsigned imask = __ballot_sync(full, init);
        if (imask) {
          const int new_leaf = __shfl_sync(full, h[H_LEAF], __ffs(static_cast<int>(imask)) - 1);
          if (new_leaf != cache_leaf && n_per_leaf >= 4) {  // more rollouts of this leaf will follow
            cache_leaf = new_leaf;
            __syncwarp();
            for (int j = lane; j < A; j += 32) cw_cache_fill(B, view, p, cache_ag, cache_meta, cache_mc, j, new_leaf);
          }
        }
      }
      __syncwarp();  // (with BRM_CW_PHASE_SYNC the vote above was the barrier)
      // ---------------------------------------------------------- S2: recycle items
      for (int i0 = 0; i0 < nR; i0 += 32) {
        const int i = i0 + lane;
        bool p1 = false, p2 = false;
        uint16_t e = 0;
        if (i < nR) {
          e = listR[i];
          const int r = e >> 4, j = e & 15;
          const uint8_t* hs = cw_hstate(p, r);
          const uint8_t rf = hs[HS_RFLAG];
          if (rf & 1) finalize(r, j);
          if (rf & 2) {
            const int leaf = cw_hdr(p, r)[H_LEAF];
            if (leaf == cache_leaf)
              cw_init_from_cache(p, cache_ag, cache_meta, cache_mc, r, j);
            else
              cw_init_item(B, p, view, r, j, leaf);
            if (hs[HS_LIVE]) {
              const uint8_t f = cw_meta(p, cw_slot(p, r, j))->flags;
              p1 = cw_in_list1<MODE>(f);
              p2 = j < M && cw_in_list2<MODE>(f, j);
              if (p1) atomicAdd(&cw_hdr(p, r)[H_NSTEPS], 1);
            }
          }
        }
        if (MODE != MODE_RULES && MODE != MODE_QUERY) cw_push(list1, n1, p1, e);
        cw_push(list2, n2, p2, e);
      }
      if (kPhaseSync) {
        // (groups of warps meet at their own named barrier: waiting for the slowest of 13 warps
        // costs more than waiting for the slowest of a few)
        constexpr int kGroup = BRM_CW_SYNC_GROUP;
        constexpr int kWarps = NT / 32;
        const int g = wid / kGroup;
        const int g_warps = (g + 1) * kGroup <= kWarps ? kGroup : kWarps - g * kGroup;
        unsigned any;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(any)
            : "r"(has_work ? 1u : 0u), "r"(1 + g), "r"(g_warps * 32)
            : "memory");
        if (!any) break;
      } else {
        __syncwarp();
      }
      // ---------------------------------------------------------- phase 1
      if (MODE == MODE_RULES) {
        for (int i = lane; i < n2; i += 32) {
          const uint16_t e = list2[i];
          cw_shape_bits_item(B, p, e >> 4, e & 15);
        }
      } else if (MODE != MODE_QUERY) {
        int nX = 0;
        for (int i0 = 0; i0 < n1; i0 += 32) {
          const int i = i0 + lane;
          uint16_t e = 0;
          bool pending = false;
          if (i < n1) {
            e = list1[i];
            pending = cw_phase1_item(B, p, ctx, e >> 4, e & 15);
          }
          cw_push(listR, nX, pending, e);
        }
        __syncwarp();
        // phase 1b: the exact drivable-area test for the few agents the grid could not decide
        for (int i = lane; i < nX; i += 32) {
          const uint16_t e = listR[i];
          cw_exact_shape_item(B, p, e >> 4, e & 15);
        }
      }
      __syncwarp();
      // ---------------------------------------------------------- phase 2
      for (int i = lane; i < n2; i += 32) {
        const uint16_t e = list2[i];
        const int r = e >> 4, a = e & 15;
        if (MODE == MODE_QUERY) {
          if (args.do_check_terminal) cw_hstate(p, r)[HS_DONE] = cw_check_terminal(B, p, r) ? 1 : 0;
        } else {
          CwEval ev;
          cw_phase2_item(B, p, r, a, MODE == MODE_RULES, ev);
          if (MODE == MODE_REPLAY) {
            const CwHdrRef h = cw_hdr(p, r);
            const int64_t bt = static_cast<int64_t>(h[H_LEAF]) * args.T + h[H_K];
            for (int v = 0; v < V; ++v) args.tr.rewards_out[(bt * M + a) * V + v] = ev.r[v];
            for (int w = 0; w < 4; ++w) args.tr.labels_out[(bt * M + a) * 4 + w] = ev.w[w];
          }
        }
      }
      __syncwarp();
      if (MODE == MODE_REPLAY) {
        // the trace of this step: every agent of every rollout that executed it
        for (int s = lane; s < P * A; s += 32) {
          const int j = s / P, r = s - j * P;  // agent-major slots (cw_slot)
          const uint8_t* hs = cw_hstate(p, r);
          if (!hs[HS_BUSY] || !hs[HS_LIVE]) continue;
          const CwHdrRef h = cw_hdr(p, r);
          const int k = h[H_K];
          const int64_t bt = static_cast<int64_t>(h[H_LEAF]) * args.T + k;
          const double* rec = cw_ag(p, s);
          const CwMeta mt = *cw_meta(p, s);
          double* o = args.tr.agents_out + (bt * A + j) * 4;
          o[0] = rec[AG_X]; o[1] = rec[AG_Y]; o[2] = rec[AG_TH]; o[3] = rec[AG_V];
          args.tr.flags_out[bt * A + j] = mt.flags;
          double* fo = args.tr.frenet_out + (bt * A + j) * 3;
          fo[0] = static_cast<double>(mt.lane); fo[1] = rec[AG_S]; fo[2] = rec[AG_LAT];
          if (j == 0) args.tr.terminal_out[bt] = hs[HS_DONE];
          if (j < M) {
            const int sm = cw_mslot(p, r, j);
            uint8_t* viol = cw_mc_viol(p, sm);
            const uint32_t* qw = cw_mc_qw(p, sm);
            for (int sl = 0; sl < R; ++sl) {
              const uint32_t prev = k > 0 ? args.tr.viol_out[((bt - 1) * M + j) * R + sl] : 0u;
              args.tr.viol_out[(bt * M + j) * R + sl] = prev + viol[sl];
              viol[sl] = 0;
              args.tr.q_out[(bt * M + j) * R + sl] = static_cast<uint8_t>(cw_q_get(qw, sl));
            }
          }
        }
        __syncwarp();
      }
    }
    if (single) break;
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

template <int MODE, int NT, int CTAS>
int launch_config(brm_engine* e, const CwLaunchCtx& ctx, CwArgs& a, int64_t units) {
  constexpr int kThreads = NT, kWarps = NT / 32;
  auto kernel = cw_pool_kernel<MODE, NT, CTAS>;
  // the opt-in shared-memory size is a per-device attribute of the function: set it on every
  // launch (an engine per GPU in one process is a supported use)
  BRM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(ctx.smem_bytes)));
  int per_sm = 0;
  BRM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, ctx.smem_bytes));
  BRM_REQUIRE(per_sm >= 1, "continuous-world kernel does not fit on an SM (%zu bytes of shared memory)",
              ctx.smem_bytes);
  int64_t grid = static_cast<int64_t>(e->sm_count) * per_sm;
  // when the launch cannot fill the GPU, spread it: a warp claims at least 4 rollouts at a time, so
  // 4 per warp is the finest useful grain -- a small call (a host tree's round) then runs on many
  // SMs with a few rollouts per warp instead of on a few SMs with full pools (8 leaves x 16 rollouts
  // 230 -> 174 us, 16384 rollouts 475 -> 210 us); a launch that fills the GPU is not affected
  const int64_t per_cta = 4 * kWarps;
  // (also when the CTAs walk over scenarios: the CTAs without a leaf of their own join the others')
  const int64_t need = (units + per_cta - 1) / per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  a.n_warps_total = static_cast<int32_t>(grid) * kWarps;
  kernel<<<static_cast<int>(grid), kThreads, ctx.smem_bytes, e->stream>>>(a);
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}

// two shapes of the same kernel: one large CTA per SM (one copy of the scenario blob, the largest
// pools) when every rollout lives in one scenario; three small CTAs per SM when the CTAs walk
// over scenarios (a CTA's warps share the rollouts of one leaf at a time)
template <int MODE>
int launch_mode(brm_engine* e, const CwLaunchCtx& ctx, CwArgs& a, int64_t units) {
  if (ctx.threads == kCwThreadsLarge) return launch_config<MODE, kCwThreadsLarge, 1>(e, ctx, a, units);
  return launch_config<MODE, kCwThreadsSmall, kCwCtasSmall>(e, ctx, a, units);
}

void fill_common(const CwLaunchCtx& ctx, CwArgs& a) {
  a.blobs = ctx.blobs;
  a.blob_bytes = ctx.blob_bytes;
  a.blob_stride = ctx.blob_stride;
  a.blob_cap = ctx.blob_cap;
  a.A = ctx.A; a.M = ctx.M; a.V = ctx.V; a.R = ctx.R; a.P = ctx.P;
  a.warp_bytes = static_cast<int32_t>(ctx.warp_bytes);
  a.n_warps_total = 1;
  a.target_items = 32 * 1024;  // batch modes: every free slot takes a state
}

}  // namespace

int cw_launch_rollout(brm_engine* e, const CwLaunchCtx& ctx, int n_scenarios, const brm_cw_states& leaves,
                      const brm_rollout_params& rp, const brm_cw_rollout_out& out, long long* zeroed_ws) {
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
  a.target_items = ctx.target_items;
  a.total = static_cast<int64_t>(leaves.n) * rp.rollouts_per_leaf;
  BRM_REQUIRE(a.total + 4 * 32 < (1ll << 31), "brm_cw_rollout: %lld rollouts in one call (limit 2^31)",
              static_cast<long long>(a.total));
  const size_t fx_count = static_cast<size_t>(leaves.n) * M * V;
  int rc = BRM_OK;
  if (zeroed_ws) {
    // the caller cleared the workspace and the sum outputs together with one memset
    a.returns_fx = zeroed_ws;
  } else {
    const size_t ws = cw_rollout_ws_words(static_cast<size_t>(leaves.n), M, V) * sizeof(long long);
    rc = e->ensure_ws(ws);
    if (rc != BRM_OK) return rc;
    a.returns_fx = static_cast<long long*>(e->d_ws);
    BRM_CUDA_CHECK(cudaMemsetAsync(a.returns_fx, 0, ws, e->stream));
    if (out.viol_sum)
      BRM_CUDA_CHECK(cudaMemsetAsync(out.viol_sum, 0, static_cast<size_t>(leaves.n) * M * R * sizeof(uint64_t),
                                     e->stream));
    if (out.agent_steps)
      BRM_CUDA_CHECK(cudaMemsetAsync(out.agent_steps, 0, static_cast<size_t>(leaves.n) * sizeof(uint64_t),
                                     e->stream));
  }
  a.work_counter = reinterpret_cast<unsigned long long*>(a.returns_fx + fx_count);
  a.leaf_next = reinterpret_cast<int32_t*>(a.work_counter + 1);
  e->time_begin();
  rc = launch_mode<MODE_ROLLOUT>(e, ctx, a, a.total);
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
