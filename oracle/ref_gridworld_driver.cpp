// TEST INFRASTRUCTURE -- driver around the reference's OWN GridWorld sources.
//
// This file is ours; everything it drives is the unmodified reference code compiled
// from /root/reference (see oracle/build_ref.sh): GridWorldState::Execute / Step /
// SetCollisionPositions / GetActionCost / GetTerminalReward / GetNumActions
// (gridworld/grid_world_state.cpp:45-226) and the four label evaluators
// (gridworld/label_evaluator/*.cpp).  Scenario construction mirrors
// gridworld/base_test_env.cpp:104-140 (MakeLabels, GetAutomataVec).
// The output library oracle/_ref/libgw_ref.so is git-ignored.
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "gridworld/grid_world_state.hpp"
#include "gridworld/label_evaluator/evaluator_label_at_position.h"
#include "gridworld/label_evaluator/evaluator_label_collision.h"
#include "gridworld/label_evaluator/evaluator_label_ego_range.h"
#include "gridworld/label_evaluator/evaluator_label_in_direct_front.h"
#include "gridworld/label_evaluator/evaluator_label_in_range.h"
#include "oracle/gw_oracle.h"
#include "oracle/philox.h"

namespace {

const char* kLabelNames[4] = {"collision", "merged_e", "merged_x", "in_direct_front_x"};

struct RefEnv {
  GridWorldStateParameter p;
  std::vector<std::shared_ptr<EvaluatorLabelBase<World>>> labels;
  std::vector<RuleMonitor::RuleMonitorSPtr> rules;
  int A, V, R;
};

int RulesPerAgent(const gwo_config* cfg) {
  int r = 0;
  for (int k = 0; k < cfg->n_rules; ++k) {
    bool specific = false;
    for (int a = 0; a < cfg->rules[k].n_aps; ++a) specific |= cfg->rules[k].ap_label[a] >= 2;
    r += specific ? cfg->n_agents - 1 : 1;
  }
  return r;
}

RefEnv MakeEnv(const gwo_config* cfg) {
  RefEnv e;
  e.A = cfg->n_agents;
  e.V = cfg->reward_vec_size;
  e.R = RulesPerAgent(cfg);
  e.p.num_other_agents = cfg->n_agents - 1;
  e.p.state_x_length = cfg->state_x_length;
  e.p.ego_goal_reached_position = cfg->ego_goal_reached_position;
  e.p.terminal_depth_ = cfg->terminal_depth;
  e.p.merging_point = cfg->merging_point;
  e.p.speed_deviation_weight = cfg->speed_deviation_weight;
  e.p.acceleration_weight = cfg->acceleration_weight;
  e.p.potential_weight = cfg->potential_weight;
  e.p.reward_vec_size = cfg->reward_vec_size;
  e.p.action_map.assign(cfg->action_map, cfg->action_map + cfg->n_actions);
  // same four evaluators, same order, as gridworld/base_test_env.cpp:104-117
  e.labels.emplace_back(
      std::make_shared<EvaluatorLabelCollision>(kLabelNames[0], e.p.merging_point));
  e.labels.emplace_back(std::make_shared<EvaluatorLabelEgoRange>(
      kLabelNames[1], e.p.merging_point, e.p.state_x_length));
  e.labels.emplace_back(std::make_shared<EvaluatorLabelInRange>(
      kLabelNames[2], e.p.merging_point, e.p.state_x_length));
  e.labels.emplace_back(std::make_shared<EvaluatorLabelInDirectFront>(kLabelNames[3]));
  for (int k = 0; k < cfg->n_rules; ++k) {
    const gwo_rule& r = cfg->rules[k];
    std::vector<ltl::TableAp> aps;
    for (int a = 0; a < r.n_aps; ++a)
      aps.push_back({kLabelNames[r.ap_label[a]], r.ap_label[a] >= 2});
    const int stride = 1 << r.n_aps;
    std::vector<uint8_t> table(r.trans, r.trans + r.n_states * stride);
    std::vector<uint8_t> acc(r.accepting, r.accepting + r.n_states);
    e.rules.push_back(RuleMonitor::MakeTableRule(aps, r.n_states, table, acc, r.weight,
                                                 r.priority, r.init_state,
                                                 r.reset_on_violation != 0));
  }
  return e;
}

std::shared_ptr<GridWorldState> MakeState(const RefEnv& e, const int32_t* agents0,
                                          const uint8_t* term0, int depth0, const uint8_t* q0) {
  std::vector<AgentState> as;
  for (int i = 0; i < e.A; ++i)
    as.emplace_back(i, agents0[4 * i + 0], agents0[4 * i + 1], agents0[4 * i + 2],
                    agents0[4 * i + 3]);
  // rule states as in gridworld/base_test_env.cpp:122-140
  RuleStateMap aut_v(e.A);
  for (int i = 0; i < e.A; ++i) {
    std::vector<int> others;
    for (int j = 0; j < e.A; ++j)
      if (j != i) others.push_back(j);
    int slot = 0;
    for (size_t k = 0; k < e.rules.size(); ++k) {
      auto rs = e.rules[k]->MakeRuleState(others, {});
      for (auto& s : rs) {
        if (q0) s.current_state_ = q0[i * e.R + slot];
        ++slot;
        aut_v[i].insert({static_cast<Rule>(k), s});
      }
    }
  }
  std::vector<bool> term(e.A);
  for (int i = 0; i < e.A; ++i) term[i] = term0[i] != 0;
  return std::make_shared<GridWorldState>(as, term[0], aut_v, e.labels, e.p, depth0, term);
}

uint32_t LabelMask(const RefEnv& e, const GridWorldState& s, int agent) {
  EvaluationMap m = s.GetAgentLabels(agent);
  uint32_t bits = 0;
  auto get = [&](const Label& l) {
    auto it = m.find(l);
    return it != m.end() && it->second;
  };
  if (get(Label(kLabelNames[0]))) bits |= 1u;
  if (get(Label(kLabelNames[1]))) bits |= 2u;
  for (int j = 0; j < e.A; ++j) {
    if (j == agent) continue;
    if (get(Label(kLabelNames[2], j))) bits |= 1u << (2 + j);
    if (get(Label(kLabelNames[3], j))) bits |= 1u << (2 + e.A + j);
  }
  return bits;
}

void RolloutRange(const RefEnv& e, const std::shared_ptr<GridWorldState>& root, uint64_t seed,
                  uint64_t begin, uint64_t n, int max_steps, double gamma, int first_exp,
                  double* returns_sum, uint64_t* viol_sum, uint64_t* agent_steps) {
  std::vector<double> acc(e.A * e.V);
  for (uint64_t r = begin; r < begin + n; ++r) {
    std::shared_ptr<GridWorldState> s = root;
    std::fill(acc.begin(), acc.end(), 0.0);
    double disc = first_exp ? gamma : 1.0;
    int k = 0;
    while (!s->IsTerminal() && k < max_steps) {
      JointAction ja(e.A);
      for (int a = 0; a < e.A; ++a) {
        const uint32_t na = static_cast<uint32_t>(s->GetNumActions(a));
        ja[a] = oracle_action_draw(seed, r, k, a, na);
        if (na > 1) ++*agent_steps;
      }
      std::vector<Reward> rew(e.A, Reward::Zero(e.V));
      s = s->Execute(ja, rew);
      for (int a = 0; a < e.A; ++a)
        for (int v = 0; v < e.V; ++v) acc[a * e.V + v] += disc * rew[a](v);
      disc *= gamma;
      ++k;
    }
    std::vector<Reward> tr = s->GetTerminalReward();
    for (int a = 0; a < e.A; ++a) {
      for (int v = 0; v < e.V; ++v) returns_sum[a * e.V + v] += acc[a * e.V + v] + disc * tr[a](v);
      int slot = 0;
      for (const auto& kv : s->GetRuleStateMap()[a]) viol_sum[a * e.R + slot++] += kv.second.violated_;
    }
  }
}

}  // namespace

extern "C" {

int32_t gwref_rules_per_agent(const gwo_config* cfg) { return RulesPerAgent(cfg); }

int32_t gwref_replay(GWO_REPLAY_ARGS) {
  RefEnv e = MakeEnv(cfg);
  std::shared_ptr<GridWorldState> s = MakeState(e, agents0, term0, depth0, q0);
  int t = 0;
  for (; t < T && !s->IsTerminal(); ++t) {
    JointAction ja(e.A);
    for (int a = 0; a < e.A; ++a) ja[a] = actions[t * e.A + a];
    std::vector<Reward> rew(e.A, Reward::Zero(e.V));
    s = s->Execute(ja, rew);
    const auto& as = s->GetAgentStates();
    for (int a = 0; a < e.A; ++a) {
      int32_t* o = agents_out + (static_cast<size_t>(t) * e.A + a) * 4;
      o[0] = as[a].x_pos;
      o[1] = as[a].last_action;
      o[2] = as[a].lane;
      o[3] = as[a].init_lane;
      term_out[t * e.A + a] = s->GetNumActions(a) == 1 ? 1 : 0;
      int slot = 0;
      for (const auto& kv : s->GetRuleStateMap()[a]) {
        q_out[(static_cast<size_t>(t) * e.A + a) * e.R + slot] =
            static_cast<uint8_t>(kv.second.current_state_);
        viol_out[(static_cast<size_t>(t) * e.A + a) * e.R + slot] = kv.second.violated_;
        ++slot;
      }
      for (int v = 0; v < e.V; ++v)
        rewards_out[(static_cast<size_t>(t) * e.A + a) * e.V + v] = rew[a](v);
      labels_out[t * e.A + a] = LabelMask(e, *s, a);
    }
  }
  std::vector<Reward> tr = s->GetTerminalReward();
  for (int a = 0; a < e.A; ++a)
    for (int v = 0; v < e.V; ++v) term_reward_out[a * e.V + v] = tr[a](v);
  return t;
}

int32_t gwref_rollout(GWO_ROLLOUT_ARGS) {
  RefEnv e = MakeEnv(cfg);
  std::shared_ptr<GridWorldState> root = MakeState(e, agents0, term0, depth0, q0);
  const int P = n_threads < 1 ? 1 : n_threads;
  std::vector<std::vector<double>> rs(P, std::vector<double>(e.A * e.V, 0.0));
  std::vector<std::vector<uint64_t>> vs(P, std::vector<uint64_t>(e.A * e.R, 0));
  std::vector<uint64_t> steps(P, 0);
  std::vector<std::thread> th;
  for (int p = 0; p < P; ++p) {
    const uint64_t b = rollout_begin + n_rollouts * p / P;
    const uint64_t en = rollout_begin + n_rollouts * (p + 1) / P;
    th.emplace_back([&, p, b, en]() {
      // each worker owns a private copy of the scenario (the reference is not thread-safe)
      RefEnv le = MakeEnv(cfg);
      std::shared_ptr<GridWorldState> lroot = MakeState(le, agents0, term0, depth0, q0);
      // thread-private accumulators (no false sharing between workers), published at the end
      std::vector<double> lr(le.A * le.V, 0.0);
      std::vector<uint64_t> lv(le.A * le.R, 0);
      uint64_t ls = 0;
      RolloutRange(le, lroot, seed, b, en - b, max_steps, gamma, first_discount_exp, lr.data(),
                   lv.data(), &ls);
      rs[p] = lr;
      vs[p] = lv;
      steps[p] = ls;
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < e.A * e.V; ++i) {
    returns_sum[i] = 0;
    for (int p = 0; p < P; ++p) returns_sum[i] += rs[p][i];
  }
  for (int i = 0; i < e.A * e.R; ++i) {
    viol_sum[i] = 0;
    for (int p = 0; p < P; ++p) viol_sum[i] += vs[p][i];
  }
  *agent_steps = 0;
  for (int p = 0; p < P; ++p) *agent_steps += steps[p];
  return 0;
}

int32_t gwref_at_position(int32_t x, int32_t last_action, int32_t position) {
  EvaluatorLabelAtPosition ev("position", position);
  AgentState a(0, x, last_action, 1, 1);
  World w(a, {});
  return ev.evaluate(w)[0].second ? 1 : 0;
}

int32_t gwref_label_collision(int32_t merging_point, int32_t ex, int32_t elast, int32_t elane,
                              int32_t ox, int32_t olast, int32_t olane) {
  EvaluatorLabelCollision ev("collision", merging_point);
  AgentState e(0, ex, elast, elane, elane);
  AgentState o(1, ox, olast, olane, olane);
  World w(e, {o});
  return ev.evaluate(w)[0].second ? 1 : 0;
}

}  // extern "C"
