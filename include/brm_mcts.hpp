// brm_mcts.hpp -- a small host-side multi-agent MCTS that drives the device engine from C++
// (INTEGRATION.md level 2): the tree stays on the host, one round = several descents kept apart
// by a virtual loss, their new nodes created with the StateInterface members of brm_state.hpp
// and evaluated by ONE rollout launch for all of them (brm::*RolloutEvaluator, the replacement
// of the mvmcts::RandomHeuristic call of an iteration, src/behavior_rules_mcts.hpp:188-189,232).
//
// Selection is decoupled UCT: every agent keeps per-action visit counts and value-vector sums
// and picks its action on its own (untried first, then UCB on values normalised with the
// return bounds); value vectors are compared in thresholded lexicographic order (at objective
// k, values at or above THRESHOLD[k] count as equally good).  The reference's statistics
// classes (mvmcts::ThresUCTStatistic, un-vendored lexmamcts@24ab42ca) are not available, so
// tie-breaking and the exploration term are this header's own ("parity unpinned"); what the
// reference pins are behaviours (test/grid_world_threshold_test.cpp:40-71), reproduced by
// tests/cpp/mcts_test.cpp.  Same algorithm as planner-rules-mcts_b200/mcts.py.
//
// Header-only, C++17.  State must offer Execute / IsTerminal / GetNumActions /
// GetTerminalReward / GetAgentIdx (brm::GridWorldState, brm::MvmctsState); Evaluator must offer
// Evaluate(leaves, rollout_base) -> [leaf][agent] mean return vectors.
#ifndef BRM_MCTS_HPP_
#define BRM_MCTS_HPP_

#include <cmath>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "brm_state.hpp"

namespace brm {

// the mvmcts::MvmctsParameters fields the reference sets (src/util.cpp:23-89,
// gridworld/base_test_env.cpp:45-77)
struct MctsParameters {
  double DISCOUNT_FACTOR = 0.9;
  std::vector<double> EXPLORATION_CONSTANT;  // per objective
  std::vector<double> LOWER_BOUND, UPPER_BOUND, THRESHOLD;
  int batch_size = 8;  // descents per round (1 = the reference's loop: one expansion per iteration)
  explicit MctsParameters(int V)
      : EXPLORATION_CONSTANT(V, 0.7), LOWER_BOUND(V, -1.0), UPPER_BOUND(V, 0.0), THRESHOLD(V, 1e300) {
    LOWER_BOUND[V - 1] = -1000.0;  // src/util.cpp:58-59
  }
};

template <class State, class Evaluator>
class Mcts {
 public:
  typedef std::shared_ptr<State> StatePtr;
  typedef std::vector<Reward> JointReward;  // [agent][objective]

  Mcts(MctsParameters params, std::shared_ptr<Evaluator> evaluator, uint64_t rollout_base = 0)
      : p_(std::move(params)), eval_(std::move(evaluator)), next_rollout_(rollout_base) {}

  // Mvmcts::Search(state, max_iterations) (src/behavior_rules_mcts.hpp:232)
  void Search(const StatePtr& state, int max_iterations, int rollouts_per_expansion) {
    root_ = MakeNode(state);
    if (root_->terminal) throw Error(BRM_ERR_TERMINAL, "Current state is terminal! Aborting!");  // :208-209
    const size_t M = root_->N.size(), V = p_.LOWER_BOUND.size();
    const double gamma = p_.DISCOUNT_FACTOR;
    int done = 0;
    while (done < max_iterations) {
      const int k = std::min(p_.batch_size, max_iterations - done);
      struct Step { Node* node; JointAction ja; const JointReward* r; };
      struct Descent { std::vector<Step> path; Node* term; int exp; };
      std::vector<Descent> descents;
      std::vector<std::pair<Node*, JointAction>> expansions;
      std::map<std::pair<Node*, JointAction>, int> index;
      for (int d = 0; d < k; ++d) {
        Node* node = root_.get();
        Descent ds;
        ds.term = nullptr;
        ds.exp = -1;
        for (;;) {
          if (node->terminal) {
            ds.term = node;
            break;
          }
          const JointAction ja = Select(*node);
          for (size_t a = 0; a < M; ++a) {  // virtual loss: counted at once, valued at the lower bound
            node->N[a][ja[a]] += 1.0;
            for (size_t v = 0; v < V; ++v) node->Q[a][ja[a]][v] += p_.LOWER_BOUND[v];
          }
          auto it = node->children.find(ja);
          if (it != node->children.end()) {
            ds.path.push_back(Step{node, ja, &it->second.second});
            node = it->second.first.get();
            continue;
          }
          const auto key = std::make_pair(node, ja);
          auto ex = index.find(key);
          if (ex == index.end()) {
            ex = index.emplace(key, static_cast<int>(expansions.size())).first;
            expansions.push_back(key);
          }
          ds.path.push_back(Step{node, ja, nullptr});
          ds.exp = ex->second;
          break;
        }
        descents.push_back(std::move(ds));
      }
      // the new nodes of the round: Execute per expansion, ONE rollout launch for all of them
      std::vector<Node*> fresh;
      if (!expansions.empty()) {
        std::vector<StatePtr> leaves;
        std::vector<JointReward> step_rewards;
        for (auto& e : expansions) {
          JointReward r;
          leaves.push_back(e.first->state->Execute(e.second, r));
          step_rewards.push_back(std::move(r));
        }
        const auto values = eval_->Evaluate(leaves, next_rollout_);
        next_rollout_ += static_cast<uint64_t>(leaves.size()) * rollouts_per_expansion;
        for (size_t i = 0; i < expansions.size(); ++i) {
          auto child = MakeNode(leaves[i]);
          child->value = values[i];
          fresh.push_back(child.get());
          expansions[i].first->children.emplace(expansions[i].second, std::make_pair(std::move(child), step_rewards[i]));
        }
      }
      for (const Descent& ds : descents) {
        JointReward value;
        const JointReward* r_new = nullptr;
        if (ds.term) {
          if (ds.term->term_value.empty()) ds.term->term_value = ds.term->state->GetTerminalReward();
          value = ds.term->term_value;
        } else {
          value = fresh[ds.exp]->value;
          r_new = &expansions[ds.exp].first->children.find(expansions[ds.exp].second)->second.second;
        }
        for (auto it = ds.path.rbegin(); it != ds.path.rend(); ++it) {
          const JointReward& r = it->r ? *it->r : *r_new;
          for (size_t a = 0; a < M; ++a)
            for (size_t v = 0; v < V; ++v) {
              value[a][v] = r[a][v] + gamma * value[a][v];
              it->node->Q[a][it->ja[a]][v] += value[a][v] - p_.LOWER_BOUND[v];  // replaces the virtual loss
            }
        }
        ++iterations_;
      }
      done += k;
    }
  }

  // Mvmcts::ReturnBestAction(): per agent the root action with the best mean return vector
  JointAction ReturnBestAction() const {
    const size_t M = root_->N.size();
    JointAction best(M, 0);
    for (size_t a = 0; a < M; ++a) {
      int b = -1;
      std::vector<double> bm;
      for (size_t k = 0; k < root_->N[a].size(); ++k) {
        if (root_->N[a][k] <= 0.0) continue;
        std::vector<double> m = root_->Q[a][k];
        for (double& x : m) x /= root_->N[a][k];
        if (b < 0 || LexBetter(m, bm, p_.THRESHOLD)) {
          b = static_cast<int>(k);
          bm = m;
        }
      }
      best[a] = b < 0 ? 0 : static_cast<ActionIdx>(b);
    }
    return best;
  }

  // GetEgoIntNode().GetExpectedRewards() (gridworld/grid_world_env.h:38): root action -> mean return
  std::vector<Reward> GetExpectedRewards(AgentIdx agent = 0) const {
    std::vector<Reward> out;
    for (size_t k = 0; k < root_->N[agent].size(); ++k) {
      Reward m = root_->Q[agent][k];
      for (double& x : m) x = root_->N[agent][k] > 0.0 ? x / root_->N[agent][k] : std::nan("");
      out.push_back(m);
    }
    return out;
  }

  // packed root statistics [agent][action][1 + V] (visits, value sums): what brm_allreduce_root_stats sums
  std::vector<double> GetRootStats(int n_actions) const {
    const size_t M = root_->N.size(), V = p_.LOWER_BOUND.size();
    std::vector<double> s(M * n_actions * (1 + V), 0.0);
    for (size_t a = 0; a < M; ++a)
      for (size_t k = 0; k < root_->N[a].size(); ++k) {
        double* o = &s[(a * n_actions + k) * (1 + V)];
        o[0] = root_->N[a][k];
        for (size_t v = 0; v < V; ++v) o[1 + v] = root_->Q[a][k][v];
      }
    return s;
  }

  int NumIterations() const { return iterations_; }

 private:
  struct Node {
    StatePtr state;
    bool terminal = false;
    std::vector<std::vector<double>> N;               // [agent][action]
    std::vector<std::vector<std::vector<double>>> Q;  // [agent][action][objective]
    JointReward value, term_value;
    std::map<JointAction, std::pair<std::unique_ptr<Node>, JointReward>> children;
  };

  std::unique_ptr<Node> MakeNode(const StatePtr& s) const {
    auto n = std::unique_ptr<Node>(new Node());
    n->state = s;
    n->terminal = s->IsTerminal();
    const size_t M = s->GetAgentIdx().size(), V = p_.LOWER_BOUND.size();
    for (size_t a = 0; a < M; ++a) {
      const size_t k = s->GetNumActions(a);
      n->N.emplace_back(k, 0.0);
      n->Q.emplace_back(k, std::vector<double>(V, 0.0));
    }
    return n;
  }

  // a better than b in thresholded lexicographic order
  static bool LexBetter(const std::vector<double>& a, const std::vector<double>& b, const std::vector<double>& thr) {
    for (size_t k = 0; k < a.size(); ++k) {
      const double x = std::min(a[k], thr[k]), y = std::min(b[k], thr[k]);
      if (std::fabs(x - y) > 1e-12) return x > y;
    }
    return false;
  }

  JointAction Select(const Node& node) const {
    const size_t M = node.N.size(), V = p_.LOWER_BOUND.size();
    JointAction ja(M, 0);
    std::vector<double> thr(V);
    for (size_t v = 0; v < V; ++v) {
      const double span = std::max(p_.UPPER_BOUND[v] - p_.LOWER_BOUND[v], 1e-9);
      thr[v] = p_.THRESHOLD[v] >= 1e299 ? 1e300 : (p_.THRESHOLD[v] - p_.LOWER_BOUND[v]) / span;
    }
    for (size_t a = 0; a < M; ++a) {
      const size_t n_act = node.N[a].size();
      bool untried = false;
      double total = 0.0;
      for (size_t k = 0; k < n_act; ++k) {
        if (node.N[a][k] == 0.0 && !untried) {
          ja[a] = k;
          untried = true;
        }
        total += node.N[a][k];
      }
      if (untried) continue;
      int best = -1;
      std::vector<double> bu;
      for (size_t k = 0; k < n_act; ++k) {
        std::vector<double> u(V);
        const double bonus = std::sqrt(2.0 * std::log(total) / node.N[a][k]);
        for (size_t v = 0; v < V; ++v) {
          const double span = std::max(p_.UPPER_BOUND[v] - p_.LOWER_BOUND[v], 1e-9);
          u[v] = (node.Q[a][k][v] / node.N[a][k] - p_.LOWER_BOUND[v]) / span + bonus * p_.EXPLORATION_CONSTANT[v];
        }
        if (best < 0 || LexBetter(u, bu, thr)) {
          best = static_cast<int>(k);
          bu = u;
        }
      }
      ja[a] = static_cast<ActionIdx>(best);
    }
    return ja;
  }

  MctsParameters p_;
  std::shared_ptr<Evaluator> eval_;
  uint64_t next_rollout_;
  std::unique_ptr<Node> root_;
  int iterations_ = 0;
};

}  // namespace brm

#endif  // BRM_MCTS_HPP_
