// include/brm_mcts.hpp on the CPU: the host tree driven by a toy state and a toy evaluator (no
// engine, no GPU), so that selection, virtual loss, back-propagation and the thresholded
// lexicographic order are tested in the `-m "not gpu"` suite.
//   mcts_logic_test            prints one line per check; exit code 0 when all hold
#include <cstdio>
#include <cstdlib>

#include "brm_mcts.hpp"

// One or two agents walk down a tree of fixed depth.  Reward vector (V = 2), both <= 0:
//   objective 0 ("rule"): -1 when the agent takes action 2
//   objective 1 ("cost"): action 0: -0.5, action 1: -0.1, action 2: 0
// Lexicographic: never violate the rule, then the cheapest action -> action 1.  With a threshold
// of -10 on objective 0 (everything counts as equally good there) the cost decides -> action 2.
struct ToyState {
  int depth, horizon, n_agents;
  ToyState(int d, int h, int n) : depth(d), horizon(h), n_agents(n) {}
  std::shared_ptr<ToyState> Execute(const brm::JointAction& ja, std::vector<brm::Reward>& r) const {
    r.assign(n_agents, brm::Reward(2, 0.0));
    for (int a = 0; a < n_agents; ++a) {
      r[a][0] = ja[a] == 2 ? -1.0 : 0.0;
      r[a][1] = ja[a] == 0 ? -0.5 : (ja[a] == 1 ? -0.1 : 0.0);
    }
    return std::make_shared<ToyState>(depth + 1, horizon, n_agents);
  }
  bool IsTerminal() const { return depth >= horizon; }
  std::size_t GetNumActions(brm::AgentIdx) const { return 3; }
  std::vector<brm::Reward> GetTerminalReward() const { return std::vector<brm::Reward>(n_agents, brm::Reward(2, 0.0)); }
  std::vector<brm::AgentIdx> GetAgentIdx() const {
    std::vector<brm::AgentIdx> v(n_agents);
    for (int a = 0; a < n_agents; ++a) v[a] = a;
    return v;
  }
};

struct ToyEvaluator {
  int calls = 0, leaves = 0;
  uint64_t last_base = 0;
  std::vector<std::vector<brm::Reward>> Evaluate(const std::vector<std::shared_ptr<ToyState>>& l, uint64_t base) {
    ++calls;
    leaves += static_cast<int>(l.size());
    last_base = base;
    return std::vector<std::vector<brm::Reward>>(l.size(), std::vector<brm::Reward>(l[0]->n_agents, brm::Reward(2, 0.0)));
  }
};

static int failures = 0;
static void check(bool ok, const char* what) {
  std::printf("%s: %s\n", ok ? "ok" : "FAIL", what);
  if (!ok) ++failures;
}

int main() {
  typedef brm::Mcts<ToyState, ToyEvaluator> Search;
  for (int batch : {1, 8}) {
    for (int n_agents : {1, 2}) {
      brm::MctsParameters mp(2);
      mp.LOWER_BOUND = {-4.0, -4.0};
      mp.UPPER_BOUND = {0.0, 0.0};
      mp.DISCOUNT_FACTOR = 0.9;
      mp.batch_size = batch;
      auto ev = std::make_shared<ToyEvaluator>();
      Search m(mp, ev, 1000);
      m.Search(std::make_shared<ToyState>(0, 4, n_agents), 600, 16);
      const std::vector<double> st = m.GetRootStats(3);
      bool visits_ok = true, best_ok = true;
      for (int a = 0; a < n_agents; ++a) {
        double visits = 0.0;
        for (int k = 0; k < 3; ++k) visits += st[(a * 3 + k) * 3];
        visits_ok = visits_ok && visits == 600.0;          // every descent is counted exactly once
        best_ok = best_ok && m.ReturnBestAction()[a] == 1;  // no violation, then the cheapest
      }
      check(visits_ok, "root visits equal the iterations (virtual losses are all replaced)");
      check(best_ok, "lexicographic order: the rule first, then the cost");
      check(m.NumIterations() == 600, "NumIterations");
      // root action 2 always pays the rule penalty of the first step; the means stay within the bounds
      const std::vector<brm::Reward> q = m.GetExpectedRewards(0);
      check(q[2][0] <= -1.0 + 1e-12 && q[1][0] > q[2][0] + 0.5 && q[1][1] < 0.0 && q[1][1] > -4.0 && q[0][1] < q[1][1],
            "expected rewards of the root actions");
      check(batch == 1 ? ev->calls == ev->leaves : ev->calls < ev->leaves,
            "one evaluator call per round: several new leaves per call when the batch is larger than 1");
      const uint64_t off = ev->last_base - 1000;
      check(batch == 1 ? off == 16ull * (ev->leaves - 1) : (off % 16 == 0 && off < 16ull * ev->leaves),
            "rollout ids advance by leaves x rollouts per expansion");
    }
  }
  {
    brm::MctsParameters mp(2);
    mp.LOWER_BOUND = {-4.0, -4.0};
    mp.UPPER_BOUND = {0.0, 0.0};
    mp.THRESHOLD = {-10.0, 1e300};  // objective 0: everything at or above -10 is equally good
    mp.batch_size = 4;
    auto ev = std::make_shared<ToyEvaluator>();
    Search m(mp, ev, 0);
    m.Search(std::make_shared<ToyState>(0, 4, 1), 600, 16);
    check(m.ReturnBestAction()[0] == 2, "thresholded order: below the threshold the next objective decides");
  }
  {
    brm::MctsParameters mp(2);
    auto ev = std::make_shared<ToyEvaluator>();
    Search m(mp, ev, 0);
    bool threw = false;
    try {
      m.Search(std::make_shared<ToyState>(4, 4, 1), 10, 16);
    } catch (const brm::Error& e) {
      threw = e.code == BRM_ERR_TERMINAL;
    }
    check(threw, "a terminal root is refused (src/behavior_rules_mcts.hpp:208-209)");
  }
  std::printf("%s\n", failures ? "EXPECTATIONS FAIL" : "expectations hold");
  return failures ? 1 : 0;
}
