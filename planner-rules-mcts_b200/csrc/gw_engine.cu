// GridWorld host API (include/brm.h, brm_gw_*): scenario upload and the batched
// StateInterface surface.  Kernels live in gw_kernels.cu.
#include "common.cuh"
#include "staging.cuh"

namespace brm {
int gw_launch_rollout(brm_engine* e, const brm_gw_states& leaves, const brm_rollout_params& rp,
                      const brm_gw_rollout_out& out, bool sums_zeroed = false);
int gw_launch_execute(brm_engine* e, const brm_gw_states& in, const uint8_t* actions,
                      const brm_gw_states& out, double* rewards);
int gw_launch_query(brm_engine* e, const brm_gw_states& s, uint8_t* num_actions,
                    uint8_t* is_terminal, double* term_reward);
int gw_launch_replay(brm_engine* e, const brm_gw_states& s0, const uint8_t* actions, int32_t T,
                     const brm_gw_trace& tr);

namespace {

int check_ready(const brm_engine* e, const char* fn) {
  if (!e) {
    set_error("%s: engine is NULL", fn);
    return BRM_ERR_INVALID;
  }
  if (!e->gw_ready) {
    set_error("%s: call brm_gw_configure first", fn);
    return BRM_ERR_NOT_CONFIGURED;
  }
  return BRM_OK;
}

int check_states(const brm_gw_states* s, const char* fn) {
  BRM_REQUIRE(s != nullptr, "%s: states is NULL", fn);
  BRM_REQUIRE(s->n > 0, "%s: empty batch (n=%d)", fn, s->n);
  BRM_REQUIRE(s->agents && s->term && s->depth && s->q, "%s: agents/term/depth/q must be set", fn);
  BRM_REQUIRE(s->mem == BRM_MEM_HOST || s->mem == BRM_MEM_DEVICE, "%s: bad mem %d", fn, s->mem);
  return BRM_OK;
}

int upload_states(Stager& st, const brm_gw_states& h, int A, int R, brm_gw_states* d) {
  const size_t n = h.n;
  *d = h;
  d->mem = BRM_MEM_DEVICE;
  BRM_TRY(st.in(h.agents, A * n * 4, &d->agents));
  BRM_TRY(st.in(h.term, n, &d->term));
  BRM_TRY(st.in(h.depth, n, &d->depth));
  BRM_TRY(st.in(h.q, A * R * n, &d->q));
  if (h.viol) BRM_TRY(st.in(h.viol, A * R * n, &d->viol));
  return BRM_OK;
}

int alloc_states(Stager& st, const brm_gw_states& like, int A, int R, brm_gw_states* d) {
  const size_t n = like.n;
  *d = like;
  d->mem = BRM_MEM_DEVICE;
  BRM_TRY(st.alloc(A * n * 4, &d->agents));
  BRM_TRY(st.alloc(n, &d->term));
  BRM_TRY(st.alloc(n, &d->depth));
  BRM_TRY(st.alloc(A * R * n, &d->q));
  if (like.viol) BRM_TRY(st.alloc(A * R * n, &d->viol, true));
  return BRM_OK;
}

int download_states(Stager& st, const brm_gw_states& d, int A, int R, brm_gw_states* h) {
  const size_t n = h->n;
  BRM_TRY(st.out(h->agents, d.agents, A * n * 4));
  BRM_TRY(st.out(h->term, d.term, n));
  BRM_TRY(st.out(h->depth, d.depth, n));
  BRM_TRY(st.out(h->q, d.q, A * R * n));
  if (h->viol) BRM_TRY(st.out(h->viol, d.viol, A * R * n));
  return BRM_OK;
}

}  // namespace
}  // namespace brm

using namespace brm;

extern "C" {

int brm_gw_configure(brm_engine* e, const brm_gw_params* p, const brm_rule* rules, int32_t n_rules) {
  BRM_REQUIRE(e && p, "brm_gw_configure: NULL argument");
  BRM_REQUIRE(p->n_agents >= 1 && p->n_agents <= BRM_GW_MAX_AGENTS,
              "brm_gw_configure: n_agents %d not in [1,%d]", p->n_agents, BRM_GW_MAX_AGENTS);
  BRM_REQUIRE(p->reward_vec_size >= 1 && p->reward_vec_size <= BRM_MAX_REWARD,
              "brm_gw_configure: reward_vec_size %d not in [1,%d]", p->reward_vec_size, BRM_MAX_REWARD);
  BRM_REQUIRE(p->n_actions >= 1 && p->n_actions <= BRM_MAX_ACTIONS,
              "brm_gw_configure: n_actions %d not in [1,%d]", p->n_actions, BRM_MAX_ACTIONS);
  BRM_REQUIRE(n_rules >= 0 && n_rules <= BRM_MAX_RULES, "brm_gw_configure: n_rules %d not in [0,%d]",
              n_rules, BRM_MAX_RULES);
  BRM_REQUIRE(n_rules == 0 || rules, "brm_gw_configure: rules is NULL");
  // built in a local copy: a reconfiguration that fails half-way (a bad rule, too many rule
  // states) must leave the engine exactly as it was -- host sizes and device tables in step
  GwBlob local;
  GwBlob& B = local;
  memset(&B, 0, sizeof(B));
  B.A = p->n_agents;
  B.V = p->reward_vec_size;
  B.n_actions = p->n_actions;
  B.merging_point = p->merging_point;
  B.x_len = p->state_x_length;
  B.terminal_depth = p->terminal_depth;
  B.n_rules = n_rules;
  for (int k = 0; k < p->n_actions; ++k) B.action_map[k] = p->action_map[k];
  B.w_speed = p->speed_deviation_weight;
  B.w_acc = p->acceleration_weight;
  for (int k = 0; k < n_rules; ++k) {
    const brm_rule& r = rules[k];
    BRM_REQUIRE(r.n_aps <= BRM_MAX_APS && r.n_states >= 1 && r.n_states <= BRM_MAX_DFA_STATES,
                "brm_gw_configure: rule %d has bad sizes (n_aps=%d n_states=%d)", k, r.n_aps, r.n_states);
    BRM_REQUIRE(r.priority < p->reward_vec_size, "brm_gw_configure: rule %d priority %d >= V", k,
                r.priority);
    BRM_REQUIRE(r.init_state < r.n_states, "brm_gw_configure: rule %d init_state out of range", k);
    for (int a = 0; a < r.n_aps; ++a) {
      BRM_REQUIRE(r.ap_label[a] <= BRM_GW_LABEL_IN_DIRECT_FRONT_X,
                  "brm_gw_configure: rule %d AP %d: unknown GridWorld label kind %d", k, a, r.ap_label[a]);
      const bool per_other = r.ap_label[a] >= BRM_GW_LABEL_MERGED_X;
      BRM_REQUIRE(per_other == (r.ap_agent_specific[a] != 0),
                  "brm_gw_configure: rule %d AP %d: label kind %d %s a #0 placeholder", k, a,
                  r.ap_label[a], per_other ? "needs" : "cannot take");
    }
    for (int s = 0; s < r.n_states * (1 << r.n_aps); ++s)
      BRM_REQUIRE(r.trans[s] == BRM_VIOLATION || r.trans[s] < r.n_states,
                  "brm_gw_configure: rule %d transition %d out of range", k, s);
    B.rules[k] = r;
    // device order of the letter bits: AP a at bit a
    for (int q = 0; q < r.n_states; ++q)
      for (int l = 0; l < (1 << r.n_aps); ++l) {
        int nl = 0;
        for (int a = 0; a < r.n_aps; ++a) nl |= ((l >> (r.n_aps - 1 - a)) & 1) << a;
        B.rules[k].trans[q * (1 << r.n_aps) + nl] = r.trans[q * (1 << r.n_aps) + l];
      }
  }
  // rule-slot layout per agent (BaseTestEnv::GetAutomataVec, gridworld/base_test_env.cpp:122-140)
  int R = 0;
  for (int i = 0; i < B.A; ++i) {
    int n = 0;
    for (int k = 0; k < n_rules; ++k) {
      bool specific = false;
      for (int a = 0; a < rules[k].n_aps; ++a) specific |= rules[k].ap_agent_specific[a] != 0;
      if (specific) {
        for (int j = 0; j < B.A; ++j)
          if (j != i) {
            BRM_REQUIRE(n < BRM_GW_MAX_SLOTS, "brm_gw_configure: more than %d rule states per agent",
                        BRM_GW_MAX_SLOTS);
            B.slot_rule[i][n] = static_cast<uint8_t>(k);
            B.slot_other[i][n] = static_cast<int8_t>(j);
            ++n;
          }
      } else {
        BRM_REQUIRE(n < BRM_GW_MAX_SLOTS, "brm_gw_configure: more than %d rule states per agent",
                    BRM_GW_MAX_SLOTS);
        B.slot_rule[i][n] = static_cast<uint8_t>(k);
        B.slot_other[i][n] = -1;
        ++n;
      }
    }
    R = n;
    for (int sl = 0; sl < n; ++sl) {
      const brm_rule& r = rules[B.slot_rule[i][sl]];
      const int other = B.slot_other[i][sl];
      uint32_t bits = 0;
      for (int a = 0; a < BRM_MAX_APS; ++a) {
        uint32_t pos = 31;
        if (a < r.n_aps) {
          const int kind = r.ap_label[a];
          pos = kind == BRM_GW_LABEL_COLLISION ? 0 : kind == BRM_GW_LABEL_MERGED_E ? 1
              : kind == BRM_GW_LABEL_MERGED_X ? 2 + other : 10 + other;
        }
        bits |= pos << (8 * a);
      }
      B.slot_bits[i][sl] = bits;
      B.slot_meta[i][sl] = static_cast<uint32_t>(B.slot_rule[i][sl]) | (static_cast<uint32_t>(r.n_aps) << 8) |
                           (static_cast<uint32_t>(r.priority) << 16) |
                           (static_cast<uint32_t>(r.init_state) << 24) | (r.on_violation ? 1u << 31 : 0u);
    }
  }
  B.R = R;
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  if (!e->d_gw_blob) BRM_CUDA_CHECK(cudaMalloc(&e->d_gw_blob, sizeof(GwBlob)));
  // the device copy is about to change: until it has, the engine is not configured
  e->gw_ready = false;
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  BRM_CUDA_CHECK(cudaMemcpyAsync(e->d_gw_blob, &B, sizeof(GwBlob), cudaMemcpyHostToDevice, e->stream));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  e->gw_blob = B;
  e->gw_ready = true;
  return BRM_OK;
}

int32_t brm_gw_rules_per_agent(const brm_engine* e) { return (e && e->gw_ready) ? e->gw_blob.R : -1; }

int brm_gw_execute(brm_engine* e, const brm_gw_states* in, const uint8_t* actions,
                   brm_gw_states* out, double* rewards) {
  BRM_TRY(check_ready(e, "brm_gw_execute"));
  BRM_TRY(check_states(in, "brm_gw_execute"));
  BRM_TRY(check_states(out, "brm_gw_execute(out)"));
  BRM_REQUIRE(actions && rewards, "brm_gw_execute: actions/rewards is NULL");
  BRM_REQUIRE(in->n == out->n && in->mem == out->mem, "brm_gw_execute: in/out batch mismatch");
  BRM_REQUIRE(in->agents != out->agents && in->q != out->q && in->term != out->term && in->depth != out->depth,
              "brm_gw_execute: in and out must not share arrays");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  const int A = e->gw_blob.A, R = e->gw_blob.R, V = e->gw_blob.V;
  const size_t n = in->n;
  if (in->mem == BRM_MEM_DEVICE) return gw_launch_execute(e, *in, actions, *out, rewards);
  // Execute on a terminal state is an error in the reference's BARK-backed state
  // (src/rules_mcts_state.cpp:59); GridWorldState does not check, neither do we.
  Stager st(e);
  brm_gw_states d_in, d_out;
  uint8_t* d_act;
  double* d_rew;
  BRM_TRY(upload_states(st, *in, A, R, &d_in));
  BRM_TRY(alloc_states(st, *out, A, R, &d_out));
  BRM_TRY(st.in(actions, A * n, &d_act));
  BRM_TRY(st.alloc(A * V * n, &d_rew));
  BRM_TRY(gw_launch_execute(e, d_in, d_act, d_out, d_rew));
  BRM_TRY(download_states(st, d_out, A, R, out));
  BRM_TRY(st.out(rewards, d_rew, A * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

static int gw_query(brm_engine* e, const brm_gw_states* s, uint8_t* num_actions,
                    uint8_t* is_terminal, double* term_reward, const char* fn) {
  BRM_TRY(check_ready(e, fn));
  BRM_TRY(check_states(s, fn));
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  const int A = e->gw_blob.A, R = e->gw_blob.R, V = e->gw_blob.V;
  const size_t n = s->n;
  if (s->mem == BRM_MEM_DEVICE) return gw_launch_query(e, *s, num_actions, is_terminal, term_reward);
  Stager st(e);
  brm_gw_states d;
  BRM_TRY(upload_states(st, *s, A, R, &d));
  uint8_t *d_na = nullptr, *d_it = nullptr;
  double* d_tr = nullptr;
  if (num_actions) BRM_TRY(st.alloc(A * n, &d_na));
  if (is_terminal) BRM_TRY(st.alloc(n, &d_it));
  if (term_reward) BRM_TRY(st.alloc(A * V * n, &d_tr));
  BRM_TRY(gw_launch_query(e, d, d_na, d_it, d_tr));
  if (num_actions) BRM_TRY(st.out(num_actions, d_na, A * n));
  if (is_terminal) BRM_TRY(st.out(is_terminal, d_it, n));
  if (term_reward) BRM_TRY(st.out(term_reward, d_tr, A * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_gw_get_num_actions(brm_engine* e, const brm_gw_states* s, uint8_t* out) {
  BRM_REQUIRE(out, "brm_gw_get_num_actions: out is NULL");
  return gw_query(e, s, out, nullptr, nullptr, "brm_gw_get_num_actions");
}
int brm_gw_is_terminal(brm_engine* e, const brm_gw_states* s, uint8_t* out) {
  BRM_REQUIRE(out, "brm_gw_is_terminal: out is NULL");
  return gw_query(e, s, nullptr, out, nullptr, "brm_gw_is_terminal");
}
int brm_gw_terminal_reward(brm_engine* e, const brm_gw_states* s, double* out) {
  BRM_REQUIRE(out, "brm_gw_terminal_reward: out is NULL");
  return gw_query(e, s, nullptr, nullptr, out, "brm_gw_terminal_reward");
}

// rollouts, optionally followed by the root statistics (first_action != NULL) in the same call
static int gw_rollout_impl(brm_engine* e, const brm_gw_states* leaves, const brm_rollout_params* rp,
                           const brm_gw_rollout_out* out, const uint8_t* first_action, int32_t n_actions,
                           double* stats) {
  BRM_TRY(check_ready(e, "brm_gw_rollout"));
  BRM_TRY(check_states(leaves, "brm_gw_rollout"));
  BRM_REQUIRE(rp && out, "brm_gw_rollout: NULL argument");
  BRM_REQUIRE(out->returns_sum || first_action, "brm_gw_rollout: returns_sum is required");
  BRM_REQUIRE(rp->rollouts_per_leaf >= 1, "brm_gw_rollout: rollouts_per_leaf must be >= 1");
  BRM_REQUIRE(rp->max_steps >= 0 && rp->max_steps <= 255, "brm_gw_rollout: max_steps %d not in [0,255]",
              rp->max_steps);
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  const int A = e->gw_blob.A, R = e->gw_blob.R, V = e->gw_blob.V;
  if (leaves->mem == BRM_MEM_DEVICE) {
    BRM_REQUIRE(out->returns_sum, "brm_gw_rollout: device batches need a returns_sum buffer");
    BRM_TRY(gw_launch_rollout(e, *leaves, *rp, *out));
    if (first_action) {
      BRM_TRY(brm::launch_root_stats(e, leaves->n, A, n_actions, V, first_action, out->returns_sum,
                                     rp->rollouts_per_leaf, stats));
      BRM_TRY(brm::comm_allreduce_device(e, stats, static_cast<size_t>(A) * n_actions * (1 + V)));
    }
    return BRM_OK;
  }
  // host buffers: one upload, one memset, the launches, one download (staging.cuh, Arena)
  const size_t L = leaves->n;
  const size_t N = L * static_cast<size_t>(rp->rollouts_per_leaf);
  const size_t st_count = static_cast<size_t>(A) * (first_action ? n_actions : 0) * (1 + V);
  Arena ar(e);
  const int i_agents = ar.in(leaves->agents, static_cast<size_t>(A) * L * 4);
  const int i_term = ar.in(leaves->term, L), i_depth = ar.in(leaves->depth, L);
  const int i_q = ar.in(leaves->q, static_cast<size_t>(A) * R * L);
  const int i_fa = first_action ? ar.in(first_action, L * A) : -1;
  const int o_ret = ar.out(out->returns_sum, L * A * V, false);   // NULL host pointer: computed, not copied back
  const int o_viol = out->viol_sum ? ar.out(out->viol_sum, L * A * R, true) : -1;
  const int o_steps = out->agent_steps ? ar.out(out->agent_steps, L, true) : -1;
  const int o_ror = out->ro_returns ? ar.out(out->ro_returns, N * A * V, false) : -1;
  const int o_rov = out->ro_viol ? ar.out(out->ro_viol, N * A * R, false) : -1;
  const int o_roq = out->ro_q ? ar.out(out->ro_q, N * A * R, false) : -1;
  const int o_roa = out->ro_agents ? ar.out(out->ro_agents, N * A * 4, false) : -1;
  const int o_ros = out->ro_steps ? ar.out(out->ro_steps, N, false) : -1;
  // the statistics are accumulated into: they travel up with the inputs and come back with the outputs
  const int i_st = first_action ? ar.in(stats, st_count) : -1;
  const int o_st = first_action ? ar.out(stats, st_count, false) : -1;
  BRM_TRY(ar.commit());
  brm_gw_states d = *leaves;
  d.mem = BRM_MEM_DEVICE;
  d.agents = ar.dev<int32_t>(i_agents);
  d.term = ar.dev<uint8_t>(i_term);
  d.depth = ar.dev<int32_t>(i_depth);
  d.q = ar.dev<uint8_t>(i_q);
  d.viol = nullptr;
  brm_gw_rollout_out o = {};
  o.returns_sum = ar.dev<double>(o_ret);
  if (o_viol >= 0) o.viol_sum = ar.dev<uint64_t>(o_viol);
  if (o_steps >= 0) o.agent_steps = ar.dev<uint64_t>(o_steps);
  if (o_ror >= 0) o.ro_returns = ar.dev<double>(o_ror);
  if (o_rov >= 0) o.ro_viol = ar.dev<uint8_t>(o_rov);
  if (o_roq >= 0) o.ro_q = ar.dev<uint8_t>(o_roq);
  if (o_roa >= 0) o.ro_agents = ar.dev<int32_t>(o_roa);
  if (o_ros >= 0) o.ro_steps = ar.dev<int32_t>(o_ros);
  BRM_TRY(gw_launch_rollout(e, d, *rp, o, true));
  if (first_action) {
    double* d_st = ar.dev<double>(o_st);
    BRM_TRY(brm::launch_root_stats(e, leaves->n, A, n_actions, V, ar.dev<uint8_t>(i_fa), o.returns_sum,
                                   rp->rollouts_per_leaf, d_st, ar.dev<double>(i_st)));
    BRM_TRY(brm::comm_allreduce_device(e, d_st, st_count));
  }
  return ar.finish();
}

int brm_gw_rollout(brm_engine* e, const brm_gw_states* leaves, const brm_rollout_params* rp,
                   const brm_gw_rollout_out* out) {
  return gw_rollout_impl(e, leaves, rp, out, nullptr, 0, nullptr);
}

int brm_gw_rollout_root_stats(brm_engine* e, const brm_gw_states* leaves, const brm_rollout_params* rp,
                              const brm_gw_rollout_out* out, const uint8_t* leaf_first_action, int32_t n_actions,
                              double* stats) {
  BRM_REQUIRE(leaf_first_action && stats && n_actions >= 1, "brm_gw_rollout_root_stats: bad arguments");
  BRM_TRY(check_ready(e, "brm_gw_rollout_root_stats"));
  BRM_REQUIRE(n_actions >= e->gw_blob.n_actions,
              "brm_gw_rollout_root_stats: n_actions %d is smaller than the %d actions of the world (statistics of "
              "the actions beyond it would be dropped)", n_actions, e->gw_blob.n_actions);
  return gw_rollout_impl(e, leaves, rp, out, leaf_first_action, n_actions, stats);
}

int brm_gw_replay(brm_engine* e, const brm_gw_states* s0, const uint8_t* actions, int32_t T,
                  const brm_gw_trace* tr) {
  BRM_TRY(check_ready(e, "brm_gw_replay"));
  BRM_TRY(check_states(s0, "brm_gw_replay"));
  BRM_REQUIRE(actions && tr && T >= 1, "brm_gw_replay: bad arguments");
  BRM_REQUIRE(tr->agents_out && tr->term_out && tr->q_out && tr->viol_out && tr->rewards_out &&
                  tr->labels_out && tr->term_reward_out && tr->steps_out,
              "brm_gw_replay: every trace array is required");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  const int A = e->gw_blob.A, R = e->gw_blob.R, V = e->gw_blob.V;
  if (s0->mem == BRM_MEM_DEVICE) return gw_launch_replay(e, *s0, actions, T, *tr);
  const size_t B = s0->n, BT = B * static_cast<size_t>(T);
  Stager st(e);
  brm_gw_states d;
  BRM_TRY(upload_states(st, *s0, A, R, &d));
  uint8_t* d_act;
  BRM_TRY(st.in(actions, BT * A, &d_act));
  brm_gw_trace o = {};
  BRM_TRY(st.alloc(BT * A * 4, &o.agents_out, true));
  BRM_TRY(st.alloc(BT, &o.term_out, true));
  BRM_TRY(st.alloc(BT * A * R, &o.q_out, true));
  BRM_TRY(st.alloc(BT * A * R, &o.viol_out, true));
  BRM_TRY(st.alloc(BT * A * V, &o.rewards_out, true));
  BRM_TRY(st.alloc(BT * A * 3, &o.labels_out, true));
  BRM_TRY(st.alloc(B * A * V, &o.term_reward_out, true));
  BRM_TRY(st.alloc(B, &o.steps_out, true));
  BRM_TRY(gw_launch_replay(e, d, d_act, T, o));
  BRM_TRY(st.out(tr->agents_out, o.agents_out, BT * A * 4));
  BRM_TRY(st.out(tr->term_out, o.term_out, BT));
  BRM_TRY(st.out(tr->q_out, o.q_out, BT * A * R));
  BRM_TRY(st.out(tr->viol_out, o.viol_out, BT * A * R));
  BRM_TRY(st.out(tr->rewards_out, o.rewards_out, BT * A * V));
  BRM_TRY(st.out(tr->labels_out, o.labels_out, BT * A * 3));
  BRM_TRY(st.out(tr->term_reward_out, o.term_reward_out, B * A * V));
  BRM_TRY(st.out(tr->steps_out, o.steps_out, B));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

}  // extern "C"
