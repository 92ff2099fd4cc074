// Continuous-world host API (include/brm.h, brm_cw_*): scenario blobs, staging and the
// batched MvmctsState surface.  Kernels live in cw_kernels.cu.
#include <algorithm>
#include <cmath>
#include <cstddef>

#include "common.cuh"
#include "cw_blob_build.h"
#include "cw_core.cuh"
#include "cw_launch.cuh"
#include "staging.cuh"

namespace brm {

struct CwEngine {
  int n_scenarios = 0;
  int max_prims = 0;  // largest number of motion primitives of an MCTS agent in any scenario
  uint8_t* d_blobs = nullptr;
  int32_t* d_blob_bytes = nullptr;
  CwLaunchCtx ctx;
};

namespace {

CwEngine* cw_of(brm_engine* e) { return static_cast<CwEngine*>(e->cw); }

int check_ready(const brm_engine* e, const char* fn) {
  if (!e) {
    set_error("%s: engine is NULL", fn);
    return BRM_ERR_INVALID;
  }
  if (!e->cw) {
    set_error("%s: call brm_cw_configure first", fn);
    return BRM_ERR_NOT_CONFIGURED;
  }
  return BRM_OK;
}

int check_states(const brm_engine* e, const brm_cw_states* s, const char* fn, bool homogeneous) {
  BRM_REQUIRE(s != nullptr, "%s: states is NULL", fn);
  BRM_REQUIRE(s->n > 0, "%s: empty batch (n=%d)", fn, s->n);
  BRM_REQUIRE(s->scenario && s->agents && s->flags && s->horizon && s->terminal && s->q,
              "%s: scenario/agents/flags/horizon/terminal/q must be set", fn);
  BRM_REQUIRE(s->mem == BRM_MEM_HOST || s->mem == BRM_MEM_DEVICE, "%s: bad mem %d", fn, s->mem);
  if (s->mem == BRM_MEM_HOST) {
    const int ns = static_cast<CwEngine*>(e->cw)->n_scenarios;
    for (int i = 0; i < s->n; ++i) {
      BRM_REQUIRE(s->scenario[i] >= 0 && s->scenario[i] < ns, "%s: state %d: scenario %d out of range [0,%d)", fn,
                  i, s->scenario[i], ns);
      if (homogeneous)
        BRM_REQUIRE(s->scenario[i] == s->scenario[0],
                    "%s: all states of one batch must live in the same scenario (state %d)", fn, i);
    }
  }
  return BRM_OK;
}

int upload_states(Stager& st, const brm_cw_states& h, int A, int M, int R, brm_cw_states* d) {
  const size_t n = h.n;
  *d = h;
  d->mem = BRM_MEM_DEVICE;
  BRM_TRY(st.in(h.scenario, n, &d->scenario));
  BRM_TRY(st.in(h.agents, A * n * 4, &d->agents));
  BRM_TRY(st.in(h.flags, A * n, &d->flags));
  BRM_TRY(st.in(h.horizon, n, &d->horizon));
  BRM_TRY(st.in(h.terminal, n, &d->terminal));
  BRM_TRY(st.in(h.q, static_cast<size_t>(M) * R * n, &d->q));
  if (h.viol) BRM_TRY(st.in(h.viol, static_cast<size_t>(M) * R * n, &d->viol));
  return BRM_OK;
}

int alloc_states(Stager& st, const brm_cw_states& like, int A, int M, int R, brm_cw_states* d) {
  const size_t n = like.n;
  *d = like;
  d->mem = BRM_MEM_DEVICE;
  BRM_TRY(st.alloc(n, &d->scenario));
  BRM_TRY(st.alloc(A * n * 4, &d->agents));
  BRM_TRY(st.alloc(A * n, &d->flags));
  BRM_TRY(st.alloc(n, &d->horizon));
  BRM_TRY(st.alloc(n, &d->terminal));
  BRM_TRY(st.alloc(static_cast<size_t>(M) * R * n, &d->q));
  if (like.viol) BRM_TRY(st.alloc(static_cast<size_t>(M) * R * n, &d->viol, true));
  return BRM_OK;
}

int download_states(Stager& st, const brm_cw_states& d, int A, int M, int R, brm_cw_states* h) {
  const size_t n = h->n;
  BRM_TRY(st.out(h->scenario, d.scenario, n));
  BRM_TRY(st.out(h->agents, d.agents, A * n * 4));
  BRM_TRY(st.out(h->flags, d.flags, A * n));
  BRM_TRY(st.out(h->horizon, d.horizon, n));
  BRM_TRY(st.out(h->terminal, d.terminal, n));
  BRM_TRY(st.out(h->q, d.q, static_cast<size_t>(M) * R * n));
  if (h->viol) BRM_TRY(st.out(h->viol, d.viol, static_cast<size_t>(M) * R * n));
  return BRM_OK;
}

}  // namespace
}  // namespace brm

using namespace brm;

extern "C" {

void brm_cw_release(brm_engine* e) {
  if (!e || !e->cw) return;
  CwEngine* c = cw_of(e);
  if (c->d_blobs) cudaFree(c->d_blobs);
  if (c->d_blob_bytes) cudaFree(c->d_blob_bytes);
  delete c;
  e->cw = nullptr;
}

int brm_cw_configure(brm_engine* e, const brm_cw_scenario* scenarios, int32_t n_scenarios) {
  BRM_REQUIRE(e && scenarios && n_scenarios >= 1, "brm_cw_configure: bad arguments");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  // everything is validated and built on the host first; the engine is only touched once no
  // check can fail any more
  std::vector<CwBlob> blobs(n_scenarios);
  std::vector<int32_t> bytes(n_scenarios);
  int A = 0, M = 0, V = 0, R = 0, cap = 0;
  for (int i = 0; i < n_scenarios; ++i) {
    int Ri = 0;
    std::string err;
    if (!cw_build_blob(scenarios[i], &blobs[i], &Ri, &err)) {
      set_error("%s (scenario %d)", err.c_str(), i);
      return BRM_ERR_INVALID;
    }
    const CwBlob& B = blobs[i];
    if (i == 0) {
      A = B.A; M = B.M; V = B.V; R = Ri;
    } else if (B.A != A || B.M != M || B.V != V || Ri != R) {
      set_error("brm_cw_configure: scenario %d differs from scenario 0 in n_agents / n_mcts_agents / "
                "reward_vector_size / rule states per agent", i);
      return BRM_ERR_INVALID;
    }
    bytes[i] = (B.total_bytes + 15) & ~15;
    if (bytes[i] > cap) cap = bytes[i];
  }
  // pool size: every warp keeps its own pool; as many rollout slots (at most 32: S1 of the
  // kernel maps a lane to a slot) as fit beside the blob in the CTA's share of the SM's shared
  // memory
  int smem_optin = 0, smem_sm = 0, reserved = 0;
  BRM_CUDA_CHECK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
  BRM_CUDA_CHECK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, e->device));
  BRM_CUDA_CHECK(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, e->device));
  // one large CTA per SM when there is one scenario (one blob per SM, the largest pools), small
  // CTAs when the CTAs walk over scenarios
  const int shapes[2][2] = {{kCwThreadsLarge, 1}, {kCwThreadsSmall, kCwCtasSmall}};
  int P = 0, threads = 0;
  size_t smem = 0, warp_bytes = 0;
  for (int sh = (n_scenarios == 1 ? 0 : 1); sh < 2 && P == 0; ++sh) {
    const int n_warps = shapes[sh][0] / 32;
    for (int ctas = shapes[sh][1]; ctas >= 1 && P == 0; --ctas) {
      const size_t budget = std::min<size_t>(smem_optin, static_cast<size_t>(smem_sm) / ctas - reserved - 256);
      for (int cand = 32; cand >= 1; --cand) {
        const size_t wb = cw_pool_bytes(cand, A, M, R, V) + cw_sched_bytes(cand, A, M, R);
        const size_t need = static_cast<size_t>(cap) + wb * n_warps;
        if (need <= budget) {
          // a pool whose agents fill whole rounds of 32 work items runs better than a slightly
          // larger one that leaves a last round with a few lanes (M: 24 slots x 4 agents = 3
          // rounds beat 26 slots by 2 %)
          const int full = (cand * A / 32) * 32 / A;
          P = (full >= 1 && full * 10 >= cand * 9) ? full : cand;
          warp_bytes = cw_pool_bytes(P, A, M, R, V) + cw_sched_bytes(P, A, M, R);
          smem = static_cast<size_t>(cap) + warp_bytes * n_warps;
          threads = shapes[sh][0];
          break;
        }
      }
      // fewer resident CTAs are better than pools that cannot fill a warp's lanes
      if (P * A < 32 && ctas > 1) P = 0;
    }
  }
  BRM_REQUIRE(P >= 1, "brm_cw_configure: the scenario (%d bytes) leaves no room for a rollout in %d bytes of "
              "shared memory", cap, smem_optin);
  // phase-1 items per warp and iteration the kernel starts out with: every agent of every slot,
  // rounded down to full rounds of 32 (the warps lower it when frozen agents leave their pool
  // unable to get there)
  int target = (P * A) / 32 * 32;
  if (target < 32) target = 32;
  uint8_t* d_blobs = nullptr;
  int32_t* d_bytes = nullptr;
  const size_t stride = sizeof(CwBlob);
  cudaError_t err = cudaMalloc(&d_blobs, stride * n_scenarios);
  if (err == cudaSuccess) err = cudaMalloc(&d_bytes, sizeof(int32_t) * n_scenarios);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(d_blobs, blobs.data(), stride * n_scenarios, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(d_bytes, bytes.data(), sizeof(int32_t) * n_scenarios, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) {
    set_error("brm_cw_configure: %s", cudaGetErrorString(err));
    if (d_blobs) cudaFree(d_blobs);
    if (d_bytes) cudaFree(d_bytes);
    return BRM_ERR_CUDA;
  }
  brm_cw_release(e);
  CwEngine* c = new CwEngine();
  c->n_scenarios = n_scenarios;
  for (int i = 0; i < n_scenarios; ++i)
    for (int a = 0; a < M; ++a) c->max_prims = std::max<int>(c->max_prims, blobs[i].n_prims[a]);
  c->d_blobs = d_blobs;
  c->d_blob_bytes = d_bytes;
  c->ctx.blobs = d_blobs;
  c->ctx.blob_bytes = d_bytes;
  c->ctx.blob_stride = stride;
  c->ctx.blob_cap = cap;
  c->ctx.A = A;
  c->ctx.M = M;
  c->ctx.V = V;
  c->ctx.R = R;
  c->ctx.P = P;
  c->ctx.threads = threads;
  c->ctx.target_items = target;
  c->ctx.warp_bytes = warp_bytes;
  c->ctx.smem_bytes = smem;
  e->cw = c;
  return BRM_OK;
}

int32_t brm_cw_rules_per_agent(const brm_engine* e) {
  return (e && e->cw) ? static_cast<CwEngine*>(e->cw)->ctx.R : -1;
}

int brm_cw_check_terminal(brm_engine* e, brm_cw_states* s) {
  BRM_TRY(check_ready(e, "brm_cw_check_terminal"));
  BRM_TRY(check_states(e, s, "brm_cw_check_terminal", true));
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.A, M = c->ctx.M, R = c->ctx.R;
  if (s->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_QUERY, *s, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, nullptr);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *s, A, M, R, &d));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_QUERY, d, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, nullptr));
  BRM_TRY(st.out(s->terminal, d.terminal, static_cast<size_t>(s->n)));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_execute(brm_engine* e, const brm_cw_states* in, const uint8_t* actions, brm_cw_states* out,
                   double* rewards) {
  BRM_TRY(check_ready(e, "brm_cw_execute"));
  BRM_TRY(check_states(e, in, "brm_cw_execute", true));
  BRM_REQUIRE(out && out->scenario && out->agents && out->flags && out->horizon && out->terminal && out->q,
              "brm_cw_execute: every array of `out` except viol is required");
  BRM_REQUIRE(actions && rewards, "brm_cw_execute: actions/rewards is NULL");
  BRM_REQUIRE(in->n == out->n && in->mem == out->mem, "brm_cw_execute: in/out batch mismatch");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.A, M = c->ctx.M, R = c->ctx.R, V = c->ctx.V;
  const size_t n = in->n;
  if (in->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_EXECUTE, *in, out, actions, rewards, nullptr, nullptr, 0, 0, nullptr);
  for (size_t i = 0; i < n; ++i)
    if (in->terminal[i]) {  // BARK_EXPECT_TRUE(!IsTerminal()), src/rules_mcts_state.cpp:59
      set_error("brm_cw_execute: state %zu is terminal", i);
      return BRM_ERR_TERMINAL;
    }
  Stager st(e);
  brm_cw_states d_in, d_out;
  uint8_t* d_act;
  double* d_rew;
  BRM_TRY(upload_states(st, *in, A, M, R, &d_in));
  BRM_TRY(alloc_states(st, *out, A, M, R, &d_out));
  BRM_TRY(st.in(actions, M * n, &d_act));
  BRM_TRY(st.alloc(M * V * n, &d_rew, true));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_EXECUTE, d_in, &d_out, d_act, d_rew, nullptr, nullptr, 0, 0, nullptr));
  BRM_TRY(download_states(st, d_out, A, M, R, out));
  BRM_TRY(st.out(rewards, d_rew, M * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_evaluate_rules(brm_engine* e, const brm_cw_states* in, brm_cw_states* out, double* rewards) {
  BRM_TRY(check_ready(e, "brm_cw_evaluate_rules"));
  BRM_TRY(check_states(e, in, "brm_cw_evaluate_rules", true));
  BRM_REQUIRE(out && out->scenario && out->agents && out->flags && out->horizon && out->terminal && out->q,
              "brm_cw_evaluate_rules: every array of `out` except viol is required");
  BRM_REQUIRE(rewards, "brm_cw_evaluate_rules: rewards is NULL");
  BRM_REQUIRE(in->n == out->n && in->mem == out->mem, "brm_cw_evaluate_rules: in/out batch mismatch");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.A, M = c->ctx.M, R = c->ctx.R, V = c->ctx.V;
  const size_t n = in->n;
  if (in->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_RULES, *in, out, nullptr, rewards, nullptr, nullptr, 0, 0, nullptr);
  Stager st(e);
  brm_cw_states d_in, d_out;
  double* d_rew;
  BRM_TRY(upload_states(st, *in, A, M, R, &d_in));
  BRM_TRY(alloc_states(st, *out, A, M, R, &d_out));
  BRM_TRY(st.alloc(M * V * n, &d_rew, true));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_RULES, d_in, &d_out, nullptr, d_rew, nullptr, nullptr, 0, 0, nullptr));
  BRM_TRY(download_states(st, d_out, A, M, R, out));
  BRM_TRY(st.out(rewards, d_rew, M * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

static int cw_query(brm_engine* e, const brm_cw_states* s, uint8_t* num_actions, double* term_reward,
                    const char* fn) {
  BRM_TRY(check_ready(e, fn));
  BRM_TRY(check_states(e, s, fn, true));
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.A, M = c->ctx.M, R = c->ctx.R, V = c->ctx.V;
  const size_t n = s->n;
  if (s->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_QUERY, *s, nullptr, nullptr, nullptr, num_actions, term_reward, 0, 0,
                           nullptr);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *s, A, M, R, &d));
  uint8_t* d_na = nullptr;
  double* d_tr = nullptr;
  if (num_actions) BRM_TRY(st.alloc(M * n, &d_na));
  if (term_reward) BRM_TRY(st.alloc(M * V * n, &d_tr));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_QUERY, d, nullptr, nullptr, nullptr, d_na, d_tr, 0, 0, nullptr));
  if (num_actions) BRM_TRY(st.out(num_actions, d_na, M * n));
  if (term_reward) BRM_TRY(st.out(term_reward, d_tr, M * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_get_num_actions(brm_engine* e, const brm_cw_states* s, uint8_t* out) {
  BRM_REQUIRE(out, "brm_cw_get_num_actions: out is NULL");
  return cw_query(e, s, out, nullptr, "brm_cw_get_num_actions");
}

int brm_cw_terminal_reward(brm_engine* e, const brm_cw_states* s, double* out) {
  BRM_REQUIRE(out, "brm_cw_terminal_reward: out is NULL");
  return cw_query(e, s, nullptr, out, "brm_cw_terminal_reward");
}

// rollouts, optionally followed by the root statistics (first_action != NULL) in the same call
static int cw_rollout_impl(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                           const brm_cw_rollout_out* out, const uint8_t* first_action, int32_t n_actions,
                           double* stats) {
  BRM_TRY(check_ready(e, "brm_cw_rollout"));
  BRM_TRY(check_states(e, leaves, "brm_cw_rollout", false));
  BRM_REQUIRE(rp && out && (out->returns_sum || first_action), "brm_cw_rollout: rp/out/returns_sum is required");
  BRM_REQUIRE(rp->rollouts_per_leaf >= 1, "brm_cw_rollout: rollouts_per_leaf must be >= 1");
  BRM_REQUIRE(rp->max_steps >= 0 && rp->max_steps <= 255, "brm_cw_rollout: max_steps %d not in [0,255]",
              rp->max_steps);
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.A, M = c->ctx.M, R = c->ctx.R, V = c->ctx.V;
  if (leaves->mem == BRM_MEM_DEVICE) {
    BRM_REQUIRE(out->returns_sum, "brm_cw_rollout: device batches need a returns_sum buffer");
    BRM_TRY(cw_launch_rollout(e, c->ctx, c->n_scenarios, *leaves, *rp, *out));
    if (first_action) {
      BRM_TRY(brm::launch_root_stats(e, leaves->n, M, n_actions, V, first_action, out->returns_sum,
                                     rp->rollouts_per_leaf, stats));
      BRM_TRY(brm::comm_allreduce_device(e, stats, static_cast<size_t>(M) * n_actions * (1 + V)));
    }
    return BRM_OK;
  }
  // host buffers: one upload, one memset, the launches, one download (staging.cuh, Arena)
  const size_t L = leaves->n, N = L * static_cast<size_t>(rp->rollouts_per_leaf);
  const size_t st_count = static_cast<size_t>(M) * (first_action ? n_actions : 0) * (1 + V);
  Arena ar(e);
  const int i_scn = ar.in(leaves->scenario, L);
  const int i_agents = ar.in(leaves->agents, static_cast<size_t>(A) * L * 4);
  const int i_flags = ar.in(leaves->flags, static_cast<size_t>(A) * L);
  const int i_hor = ar.in(leaves->horizon, L), i_term = ar.in(leaves->terminal, L);
  const int i_q = ar.in(leaves->q, static_cast<size_t>(M) * R * L);
  const int i_fa = first_action ? ar.in(first_action, L * M) : -1;
  const int i_st = first_action ? ar.in(stats, st_count) : -1;
  const int t_ws = ar.tmp_zero(cw_rollout_ws_words(L, M, V) * sizeof(long long));
  const int o_ret = ar.out(out->returns_sum, L * M * V, false);  // NULL host pointer: computed, not copied back
  const int o_viol = out->viol_sum ? ar.out(out->viol_sum, L * M * R, true) : -1;
  const int o_steps = out->agent_steps ? ar.out(out->agent_steps, L, true) : -1;
  const int o_ror = out->ro_returns ? ar.out(out->ro_returns, N * M * V, false) : -1;
  const int o_rov = out->ro_viol ? ar.out(out->ro_viol, N * M * R, false) : -1;
  const int o_roq = out->ro_q ? ar.out(out->ro_q, N * M * R, false) : -1;
  const int o_roa = out->ro_agents ? ar.out(out->ro_agents, N * A * 4, false) : -1;
  const int o_rof = out->ro_flags ? ar.out(out->ro_flags, N * A, false) : -1;
  const int o_ros = out->ro_steps ? ar.out(out->ro_steps, N, false) : -1;
  const int o_st = first_action ? ar.out(stats, st_count, false) : -1;
  BRM_TRY(ar.commit());
  brm_cw_states d = *leaves;
  d.mem = BRM_MEM_DEVICE;
  d.scenario = ar.dev<int32_t>(i_scn);
  d.agents = ar.dev<double>(i_agents);
  d.flags = ar.dev<uint8_t>(i_flags);
  d.horizon = ar.dev<int32_t>(i_hor);
  d.terminal = ar.dev<uint8_t>(i_term);
  d.q = ar.dev<uint8_t>(i_q);
  d.viol = nullptr;
  brm_cw_rollout_out o = {};
  o.returns_sum = ar.dev<double>(o_ret);
  if (o_viol >= 0) o.viol_sum = ar.dev<uint64_t>(o_viol);
  if (o_steps >= 0) o.agent_steps = ar.dev<uint64_t>(o_steps);
  if (o_ror >= 0) o.ro_returns = ar.dev<double>(o_ror);
  if (o_rov >= 0) o.ro_viol = ar.dev<uint8_t>(o_rov);
  if (o_roq >= 0) o.ro_q = ar.dev<uint8_t>(o_roq);
  if (o_roa >= 0) o.ro_agents = ar.dev<double>(o_roa);
  if (o_rof >= 0) o.ro_flags = ar.dev<uint8_t>(o_rof);
  if (o_ros >= 0) o.ro_steps = ar.dev<int32_t>(o_ros);
  BRM_TRY(cw_launch_rollout(e, c->ctx, c->n_scenarios, d, *rp, o, ar.dev<long long>(t_ws)));
  if (first_action) {
    double* d_st = ar.dev<double>(o_st);
    BRM_TRY(brm::launch_root_stats(e, leaves->n, M, n_actions, V, ar.dev<uint8_t>(i_fa), o.returns_sum,
                                   rp->rollouts_per_leaf, d_st, ar.dev<double>(i_st)));
    BRM_TRY(brm::comm_allreduce_device(e, d_st, st_count));
  }
  return ar.finish();
}

int brm_cw_rollout(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                   const brm_cw_rollout_out* out) {
  return cw_rollout_impl(e, leaves, rp, out, nullptr, 0, nullptr);
}

int brm_cw_rollout_root_stats(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                              const brm_cw_rollout_out* out, const uint8_t* leaf_first_action, int32_t n_actions,
                              double* stats) {
  BRM_REQUIRE(leaf_first_action && stats && n_actions >= 1, "brm_cw_rollout_root_stats: bad arguments");
  BRM_TRY(check_ready(e, "brm_cw_rollout_root_stats"));
  BRM_REQUIRE(n_actions >= cw_of(e)->max_prims,
              "brm_cw_rollout_root_stats: n_actions %d is smaller than the %d motion primitives an agent has "
              "(statistics of the actions beyond it would be dropped)", n_actions, cw_of(e)->max_prims);
  return cw_rollout_impl(e, leaves, rp, out, leaf_first_action, n_actions, stats);
}

int brm_cw_replay(brm_engine* e, const brm_cw_states* s0, const uint8_t* actions, int32_t T,
                  const brm_cw_trace* tr) {
  BRM_TRY(check_ready(e, "brm_cw_replay"));
  BRM_TRY(check_states(e, s0, "brm_cw_replay", true));
  BRM_REQUIRE(actions && tr && T >= 1, "brm_cw_replay: bad arguments");
  BRM_REQUIRE(tr->agents_out && tr->flags_out && tr->terminal_out && tr->q_out && tr->viol_out &&
                  tr->rewards_out && tr->labels_out && tr->frenet_out && tr->term_reward_out && tr->steps_out,
              "brm_cw_replay: every trace array is required");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.A, M = c->ctx.M, R = c->ctx.R, V = c->ctx.V;
  if (s0->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_REPLAY, *s0, nullptr, actions, nullptr, nullptr, nullptr, 0, T, tr);
  const size_t B = s0->n, BT = B * static_cast<size_t>(T);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *s0, A, M, R, &d));
  uint8_t* d_act;
  BRM_TRY(st.in(actions, BT * M, &d_act));
  brm_cw_trace o = {};
  BRM_TRY(st.alloc(BT * A * 4, &o.agents_out, true));
  BRM_TRY(st.alloc(BT * A, &o.flags_out, true));
  BRM_TRY(st.alloc(BT, &o.terminal_out, true));
  BRM_TRY(st.alloc(BT * M * R, &o.q_out, true));
  BRM_TRY(st.alloc(BT * M * R, &o.viol_out, true));
  BRM_TRY(st.alloc(BT * M * V, &o.rewards_out, true));
  BRM_TRY(st.alloc(BT * M * 4, &o.labels_out, true));
  BRM_TRY(st.alloc(BT * A * 3, &o.frenet_out, true));
  BRM_TRY(st.alloc(B * M * V, &o.term_reward_out, true));
  BRM_TRY(st.alloc(B, &o.steps_out, true));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_REPLAY, d, nullptr, d_act, nullptr, nullptr, nullptr, 0, T, &o));
  BRM_TRY(st.out(tr->agents_out, o.agents_out, BT * A * 4));
  BRM_TRY(st.out(tr->flags_out, o.flags_out, BT * A));
  BRM_TRY(st.out(tr->terminal_out, o.terminal_out, BT));
  BRM_TRY(st.out(tr->q_out, o.q_out, BT * M * R));
  BRM_TRY(st.out(tr->viol_out, o.viol_out, BT * M * R));
  BRM_TRY(st.out(tr->rewards_out, o.rewards_out, BT * M * V));
  BRM_TRY(st.out(tr->labels_out, o.labels_out, BT * M * 4));
  BRM_TRY(st.out(tr->frenet_out, o.frenet_out, BT * A * 3));
  BRM_TRY(st.out(tr->term_reward_out, o.term_reward_out, B * M * V));
  BRM_TRY(st.out(tr->steps_out, o.steps_out, B));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

}  // extern "C"
