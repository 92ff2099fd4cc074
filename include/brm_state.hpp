// brm_state.hpp -- thin C++ adapter: the reference's mvmcts::StateInterface member set
// (Clone / Execute / GetNumActions / IsTerminal / GetAgentIdx / PrintState /
// GetTerminalReward; /root/reference/src/rules_mcts_state.hpp:51-64,
// /root/reference/gridworld/grid_world_state.hpp:45-71) implemented on top of the C ABI of
// brm.h, so a host-side search templated on a state type (mvmcts::Mvmcts<S, ...>) can swap
// the device-backed states in.  Header-only, C++17, no CUDA headers needed: link libbrm.so.
//
// Single states are batches of one (useful for tree bookkeeping and tests); the rollout
// hot path goes through brm::RolloutEvaluator, which replaces the mvmcts::RandomHeuristic
// call of an iteration by ONE device launch for many leaves x many rollouts.
#ifndef BRM_STATE_HPP_
#define BRM_STATE_HPP_

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "brm.h"

namespace brm {

typedef std::size_t AgentIdx;
typedef std::size_t ActionIdx;
typedef std::vector<ActionIdx> JointAction;
typedef std::vector<double> Reward;  // mvmcts::Reward is an Eigen::VectorXd (src/rules_mcts_state.cpp:202)

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc) {
  if (rc != BRM_OK) throw Error(rc, brm_last_error());
}

// One GPU + one stream (RAII around brm_engine).
class Engine {
 public:
  explicit Engine(int device = 0, void* cuda_stream = nullptr) { check(brm_create(device, cuda_stream, &e_)); }
  ~Engine() { brm_destroy(e_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  brm_engine* get() const { return e_; }

 private:
  brm_engine* e_ = nullptr;
};

// ------------------------------------------------------------------ GridWorld
// Scenario constants shared by all states of a search (what the reference passes around as
// GridWorldStateParameter + label evaluators + rule monitors).
struct GridWorldContext {
  std::shared_ptr<Engine> engine;
  brm_gw_params params;
  std::vector<brm_rule> rules;
  int A = 0, V = 0, R = 0;

  GridWorldContext(std::shared_ptr<Engine> eng, const brm_gw_params& p, const std::vector<brm_rule>& r)
      : engine(std::move(eng)), params(p), rules(r) {
    check(brm_gw_configure(engine->get(), &params, rules.data(), static_cast<int32_t>(rules.size())));
    A = params.n_agents;
    V = params.reward_vec_size;
    R = brm_gw_rules_per_agent(engine->get());
  }
};

class GridWorldState {
 public:
  static const AgentIdx ego_agent_idx = 0;

  // agents: A x (x_pos, last_action, lane, init_lane); q: A x R automaton states
  GridWorldState(std::shared_ptr<GridWorldContext> ctx, std::vector<int32_t> agents, uint8_t term_mask,
                 int32_t depth, std::vector<uint8_t> q)
      : ctx_(std::move(ctx)), agents_(std::move(agents)), term_(term_mask), depth_(depth), q_(std::move(q)) {}

  std::shared_ptr<GridWorldState> Clone() const { return std::make_shared<GridWorldState>(*this); }

  std::shared_ptr<GridWorldState> Execute(const JointAction& joint_action, std::vector<Reward>& rewards) const {
    const int A = ctx_->A, V = ctx_->V, R = ctx_->R;
    std::vector<uint8_t> act(A);
    for (int a = 0; a < A; ++a) act[a] = static_cast<uint8_t>(joint_action.at(a));
    auto next = std::make_shared<GridWorldState>(*this);
    std::vector<double> rew(static_cast<size_t>(A) * V);
    std::vector<uint32_t> viol_in(static_cast<size_t>(A) * R, 0), viol_out(static_cast<size_t>(A) * R, 0);
    brm_gw_states in = view(const_cast<GridWorldState*>(this), viol_in.data());
    brm_gw_states out = view(next.get(), viol_out.data());
    check(brm_gw_execute(ctx_->engine->get(), &in, act.data(), &out, rew.data()));
    rewards.assign(A, Reward(V, 0.0));  // the reference resizes the caller's vector (:50)
    for (int a = 0; a < A; ++a)
      for (int v = 0; v < V; ++v) rewards[a][v] = rew[static_cast<size_t>(a) * V + v];
    return next;
  }

  std::vector<Reward> GetTerminalReward() const {
    const int A = ctx_->A, V = ctx_->V;
    std::vector<double> out(static_cast<size_t>(A) * V);
    brm_gw_states s = view(const_cast<GridWorldState*>(this), nullptr);
    check(brm_gw_terminal_reward(ctx_->engine->get(), &s, out.data()));
    std::vector<Reward> r(A, Reward(V, 0.0));
    for (int a = 0; a < A; ++a)
      for (int v = 0; v < V; ++v) r[a][v] = out[static_cast<size_t>(a) * V + v];
    return r;
  }

  ActionIdx GetNumActions(AgentIdx agent_idx) const {  // gridworld/grid_world_state.cpp:154-159
    return ((term_ >> agent_idx) & 1u) ? 1 : static_cast<ActionIdx>(ctx_->params.n_actions);
  }
  bool IsTerminal() const { return term_ & 1u; }
  std::vector<AgentIdx> GetAgentIdx() const {
    std::vector<AgentIdx> idx(ctx_->A);
    for (int a = 0; a < ctx_->A; ++a) idx[a] = a;
    return idx;
  }
  std::string PrintState() const {
    std::string s = "Ego: x=" + std::to_string(agents_[0]);
    for (int a = 1; a < ctx_->A; ++a) s += ", Ag" + std::to_string(a) + ": x=" + std::to_string(agents_[4 * a]);
    return s + "\n";
  }
  void ResetDepth() { depth_ = 0; }
  int GetEgoPos() const { return agents_[0]; }
  bool EgoGoalReached() const { return agents_[0] >= ctx_->params.ego_goal_reached_position; }
  const std::vector<int32_t>& GetAgentStates() const { return agents_; }
  const std::vector<uint8_t>& GetRuleStates() const { return q_; }
  const std::shared_ptr<GridWorldContext>& Context() const { return ctx_; }

 private:
  friend class GridWorldRolloutEvaluator;
  // a batch of one: [A][1][4] and [A][R][1] coincide with the per-state layout
  static brm_gw_states view(GridWorldState* s, uint32_t* viol) {
    brm_gw_states b;
    b.n = 1;
    b.mem = BRM_MEM_HOST;
    b.agents = s->agents_.data();
    b.term = &s->term_;
    b.depth = &s->depth_;
    b.q = s->q_.data();
    b.viol = viol;
    return b;
  }
  std::shared_ptr<GridWorldContext> ctx_;
  std::vector<int32_t> agents_;
  uint8_t term_;
  int32_t depth_;
  std::vector<uint8_t> q_;
};

// Replaces mvmcts::RandomHeuristic::GetHeuristicValues for a batch of leaves: mean
// discounted return vector per leaf and agent (call site of the original:
// Mvmcts<S,SE,SO,RandomHeuristic>, gridworld/grid_world_env.h:40).
class GridWorldRolloutEvaluator {
 public:
  GridWorldRolloutEvaluator(std::shared_ptr<GridWorldContext> ctx, const brm_rollout_params& rp)
      : ctx_(std::move(ctx)), rp_(rp) {}

  // returns [leaf][agent] -> Reward (mean over rollouts_per_leaf rollouts).  With `first_action`
  // ([leaf][agent] root action each leaf descends from) and `stats` ([agent][n_actions][1+V], visits
  // and value sums, accumulated into) the root statistics are folded in the same call
  // (brm_gw_rollout_root_stats).
  std::vector<std::vector<Reward>> Evaluate(const std::vector<std::shared_ptr<GridWorldState>>& leaves,
                                            uint64_t rollout_base,
                                            const std::vector<uint8_t>* first_action = nullptr, int n_actions = 0,
                                            std::vector<double>* stats = nullptr) {
    const int A = ctx_->A, V = ctx_->V, R = ctx_->R;
    const size_t L = leaves.size();
    std::vector<int32_t> agents(static_cast<size_t>(A) * L * 4), depth(L);
    std::vector<uint8_t> term(L), q(static_cast<size_t>(A) * R * L);
    for (size_t l = 0; l < L; ++l) {
      const GridWorldState& s = *leaves[l];
      for (int a = 0; a < A; ++a) {
        std::memcpy(&agents[(static_cast<size_t>(a) * L + l) * 4], &s.agents_[4 * a], 4 * sizeof(int32_t));
        for (int r = 0; r < R; ++r) q[(static_cast<size_t>(a) * R + r) * L + l] = s.q_[a * R + r];
      }
      term[l] = s.term_;
      depth[l] = s.depth_;
    }
    brm_gw_states b;
    b.n = static_cast<int32_t>(L);
    b.mem = BRM_MEM_HOST;
    b.agents = agents.data();
    b.term = term.data();
    b.depth = depth.data();
    b.q = q.data();
    b.viol = nullptr;
    std::vector<double> sums(L * A * V);
    brm_gw_rollout_out out;
    std::memset(&out, 0, sizeof(out));
    out.returns_sum = sums.data();
    brm_rollout_params rp = rp_;
    rp.rollout_base = rollout_base;
    if (stats)
      check(brm_gw_rollout_root_stats(ctx_->engine->get(), &b, &rp, &out, first_action->data(), n_actions,
                                      stats->data()));
    else
      check(brm_gw_rollout(ctx_->engine->get(), &b, &rp, &out));
    std::vector<std::vector<Reward>> res(L, std::vector<Reward>(A, Reward(V, 0.0)));
    for (size_t l = 0; l < L; ++l)
      for (int a = 0; a < A; ++a)
        for (int v = 0; v < V; ++v) res[l][a][v] = sums[(l * A + a) * V + v] / rp.rollouts_per_leaf;
    return res;
  }

 private:
  std::shared_ptr<GridWorldContext> ctx_;
  brm_rollout_params rp_;
};

// ------------------------------------------------------------------ continuous world
struct RulesMctsContext {
  std::shared_ptr<Engine> engine;
  int A = 0, M = 0, V = 0, R = 0;
  std::vector<int32_t> n_prims;  // per MCTS agent

  RulesMctsContext(std::shared_ptr<Engine> eng, const std::vector<brm_cw_scenario>& scenarios)
      : engine(std::move(eng)) {
    check(brm_cw_configure(engine->get(), scenarios.data(), static_cast<int32_t>(scenarios.size())));
    A = scenarios[0].n_agents;
    M = scenarios[0].n_mcts_agents;
    V = scenarios[0].params.reward_vector_size;
    R = brm_cw_rules_per_agent(engine->get());
    for (int a = 0; a < M; ++a) n_prims.push_back(scenarios[0].agents[a].n_prims);
  }
};

// MvmctsState (src/rules_mcts_state.hpp:43-90) backed by the device engine.
class MvmctsState {
 public:
  static const AgentIdx ego_agent_idx = 0;

  // agents: A x (x, y, theta, v); flags: A; q: M x R.  The constructor runs CheckTerminal
  // like the reference's (src/rules_mcts_state.cpp:52).
  MvmctsState(std::shared_ptr<RulesMctsContext> ctx, int32_t scenario, std::vector<double> agents,
              std::vector<uint8_t> flags, int32_t horizon, std::vector<uint8_t> q)
      : ctx_(std::move(ctx)), scenario_(scenario), agents_(std::move(agents)), flags_(std::move(flags)),
        horizon_(horizon), terminal_(0), q_(std::move(q)) {
    brm_cw_states s = view(this, nullptr);
    check(brm_cw_check_terminal(ctx_->engine->get(), &s));
  }

  std::shared_ptr<MvmctsState> Clone() const { return std::make_shared<MvmctsState>(*this); }

  // throws brm::Error(BRM_ERR_TERMINAL) on a terminal state (BARK_EXPECT_TRUE, :59)
  std::shared_ptr<MvmctsState> Execute(const JointAction& joint_action, std::vector<Reward>& rewards) const {
    const int M = ctx_->M, V = ctx_->V, R = ctx_->R;
    std::vector<uint8_t> act(M);
    for (int a = 0; a < M; ++a) act[a] = static_cast<uint8_t>(joint_action.at(a));
    auto next = std::make_shared<MvmctsState>(*this);  // copy: Execute fills it, incl. the terminal flag
    std::vector<double> rew(static_cast<size_t>(M) * V);
    std::vector<uint32_t> viol_out(static_cast<size_t>(M) * R, 0);
    brm_cw_states in = view(const_cast<MvmctsState*>(this), nullptr);
    brm_cw_states out = view(next.get(), viol_out.data());
    check(brm_cw_execute(ctx_->engine->get(), &in, act.data(), &out, rew.data()));
    rewards.assign(M, Reward(V, 0.0));
    for (int a = 0; a < M; ++a)
      for (int v = 0; v < V; ++v) rewards[a][v] = rew[static_cast<size_t>(a) * V + v];
    return next;
  }

  // EvaluateRules() (:186-197): transit the automata on the labels of this state; mutates the
  // rule states like the reference's and returns the rule penalties per agent
  std::vector<Reward> EvaluateRules() {
    const int M = ctx_->M, V = ctx_->V, R = ctx_->R;
    MvmctsState next(*this);
    std::vector<double> rew(static_cast<size_t>(M) * V);
    std::vector<uint32_t> viol_out(static_cast<size_t>(M) * R, 0);
    brm_cw_states in = view(this, nullptr);
    brm_cw_states out = view(&next, viol_out.data());
    check(brm_cw_evaluate_rules(ctx_->engine->get(), &in, &out, rew.data()));
    q_ = next.q_;
    std::vector<Reward> rewards(M, Reward(V, 0.0));
    for (int a = 0; a < M; ++a)
      for (int v = 0; v < V; ++v) rewards[a][v] = rew[static_cast<size_t>(a) * V + v];
    return rewards;
  }

  ActionIdx GetNumActions(AgentIdx agent_idx) const {  // :129-141
    const uint8_t f = flags_.at(agent_idx);
    return ((f & BRM_CW_PRESENT) && (f & BRM_CW_VALID)) ? static_cast<ActionIdx>(ctx_->n_prims.at(agent_idx)) : 1;
  }
  bool IsTerminal() const { return terminal_ != 0; }
  const std::vector<AgentIdx> GetAgentIdx() const {
    std::vector<AgentIdx> idx(ctx_->M);
    for (int a = 0; a < ctx_->M; ++a) idx[a] = a;
    return idx;
  }
  std::string PrintState() const { return std::string(); }  // :146
  std::vector<Reward> GetTerminalReward() const {
    const int M = ctx_->M, V = ctx_->V;
    std::vector<double> out(static_cast<size_t>(M) * V);
    brm_cw_states s = view(const_cast<MvmctsState*>(this), nullptr);
    check(brm_cw_terminal_reward(ctx_->engine->get(), &s, out.data()));
    std::vector<Reward> r(M, Reward(V, 0.0));
    for (int a = 0; a < M; ++a)
      for (int v = 0; v < V; ++v) r[a][v] = out[static_cast<size_t>(a) * V + v];
    return r;
  }
  const std::vector<double>& GetAgentStates() const { return agents_; }
  const std::vector<uint8_t>& GetMultiAgentRuleState() const { return q_; }
  const std::shared_ptr<RulesMctsContext>& Context() const { return ctx_; }

 private:
  friend class RulesMctsRolloutEvaluator;
  static brm_cw_states view(MvmctsState* s, uint32_t* viol) {
    brm_cw_states b;
    b.n = 1;
    b.mem = BRM_MEM_HOST;
    b.scenario = &s->scenario_;
    b.agents = s->agents_.data();
    b.flags = s->flags_.data();
    b.horizon = &s->horizon_;
    b.terminal = &s->terminal_;
    b.q = s->q_.data();
    b.viol = viol;
    return b;
  }
  std::shared_ptr<RulesMctsContext> ctx_;
  int32_t scenario_;
  std::vector<double> agents_;
  std::vector<uint8_t> flags_;
  int32_t horizon_;
  uint8_t terminal_;
  std::vector<uint8_t> q_;
};

// Replaces mvmcts::RandomHeuristic::GetHeuristicValues for a batch of MvmctsState leaves
// (instantiated in the reference at src/behavior_rules_mcts.hpp:188-189).
class RulesMctsRolloutEvaluator {
 public:
  RulesMctsRolloutEvaluator(std::shared_ptr<RulesMctsContext> ctx, const brm_rollout_params& rp)
      : ctx_(std::move(ctx)), rp_(rp) {}

  // returns [leaf][agent] -> Reward (mean over rollouts_per_leaf rollouts)
  std::vector<std::vector<Reward>> Evaluate(const std::vector<std::shared_ptr<MvmctsState>>& leaves,
                                            uint64_t rollout_base,
                                            const std::vector<uint8_t>* first_action = nullptr, int n_actions = 0,
                                            std::vector<double>* stats = nullptr) {
    const int A = ctx_->A, M = ctx_->M, V = ctx_->V, R = ctx_->R;
    const size_t L = leaves.size();
    std::vector<int32_t> scenario(L), horizon(L);
    std::vector<double> agents(static_cast<size_t>(A) * L * 4);
    std::vector<uint8_t> flags(static_cast<size_t>(A) * L), terminal(L), q(static_cast<size_t>(M) * R * L);
    for (size_t l = 0; l < L; ++l) {
      const MvmctsState& s = *leaves[l];
      scenario[l] = s.scenario_;
      horizon[l] = s.horizon_;
      terminal[l] = s.terminal_;
      for (int a = 0; a < A; ++a) {
        std::memcpy(&agents[(static_cast<size_t>(a) * L + l) * 4], &s.agents_[4 * a], 4 * sizeof(double));
        flags[static_cast<size_t>(a) * L + l] = s.flags_[a];
      }
      for (int a = 0; a < M; ++a)
        for (int r = 0; r < R; ++r) q[(static_cast<size_t>(a) * R + r) * L + l] = s.q_[a * R + r];
    }
    brm_cw_states b;
    b.n = static_cast<int32_t>(L);
    b.mem = BRM_MEM_HOST;
    b.scenario = scenario.data();
    b.agents = agents.data();
    b.flags = flags.data();
    b.horizon = horizon.data();
    b.terminal = terminal.data();
    b.q = q.data();
    b.viol = nullptr;
    std::vector<double> sums(L * M * V);
    brm_cw_rollout_out out;
    std::memset(&out, 0, sizeof(out));
    out.returns_sum = sums.data();
    brm_rollout_params rp = rp_;
    rp.rollout_base = rollout_base;
    if (stats)
      check(brm_cw_rollout_root_stats(ctx_->engine->get(), &b, &rp, &out, first_action->data(), n_actions,
                                      stats->data()));
    else
      check(brm_cw_rollout(ctx_->engine->get(), &b, &rp, &out));
    std::vector<std::vector<Reward>> res(L, std::vector<Reward>(M, Reward(V, 0.0)));
    for (size_t l = 0; l < L; ++l)
      for (int a = 0; a < M; ++a)
        for (int v = 0; v < V; ++v) res[l][a][v] = sums[(l * M + a) * V + v] / rp.rollouts_per_leaf;
    return res;
  }

 private:
  std::shared_ptr<RulesMctsContext> ctx_;
  brm_rollout_params rp_;
};

}  // namespace brm

#endif  // BRM_STATE_HPP_
