// Continuous-world planner decisions from C++: include/brm_mcts.hpp (host tree, leaf-parallel
// rounds) over brm::MvmctsState / brm::RulesMctsRolloutEvaluator (include/brm_state.hpp) over the
// C ABI -- the call sequence of BehaviorRulesMcts::Plan (src/behavior_rules_mcts.hpp:188-249:
// build the MvmctsState, Mvmcts::Search(max_iterations), ReturnBestAction) with no Python in the
// loop.  Pins of test/single_agent_test.cpp:229-268 (free road: accelerate; slower agent close in
// front: brake).
//   planner_test <scenario.bin> <state.bin> <expected ego action> [iterations] [batch] [rollouts]
// prints the root statistics of the ego; exit code 0 when the best action is the expected one.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "brm_mcts.hpp"

template <class T>
static std::vector<T> read_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<T> out(raw.size() / sizeof(T));
  std::memcpy(out.data(), raw.data(), out.size() * sizeof(T));
  return out;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int expected = std::atoi(argv[3]);
  const int iterations = argc > 4 ? std::atoi(argv[4]) : 1000;
  const int batch = argc > 5 ? std::atoi(argv[5]) : 8;
  const int n_roll = argc > 6 ? std::atoi(argv[6]) : 16;
  try {
    auto engine = std::make_shared<brm::Engine>(0);
    std::vector<brm_cw_scenario> scn = read_file<brm_cw_scenario>(argv[1]);
    if (scn.size() != 1) return 2;
    auto ctx = std::make_shared<brm::RulesMctsContext>(engine, scn);
    const std::vector<double> agents = read_file<double>(argv[2]);  // [agent][x, y, theta, v]
    if (static_cast<int>(agents.size()) != 4 * ctx->A) return 2;
    std::vector<uint8_t> q(ctx->M * ctx->R, 0), flags(ctx->A, BRM_CW_PRESENT | BRM_CW_VALID);
    for (int a = 0; a < ctx->M; ++a) {  // initial automaton states, one slot per other agent for agent-specific rules
      int sl = 0;
      for (int k = 0; k < scn[0].n_rules; ++k) {
        const brm_rule& r = scn[0].rules[k];
        bool specific = false;
        for (int i = 0; i < r.n_aps; ++i) specific |= r.ap_agent_specific[i] != 0;
        for (int j = 0; j < (specific ? ctx->A - 1 : 1); ++j) q[a * ctx->R + sl++] = r.init_state;
      }
    }
    auto state = std::make_shared<brm::MvmctsState>(ctx, 0, agents, flags, scn[0].params.horizon, q);

    brm::MctsParameters mp(ctx->V);  // src/util.cpp:23-89
    mp.DISCOUNT_FACTOR = scn[0].params.discount_factor;
    for (double& c : mp.EXPLORATION_CONSTANT) c = 1.42;  // :51-56
    mp.batch_size = batch;
    brm_rollout_params rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.seed = 1000;
    rp.rollouts_per_leaf = n_roll;
    rp.max_steps = 30;  // MaxNumIterationsRandomHeuristic
    rp.discount_factor = mp.DISCOUNT_FACTOR;
    rp.first_discount_exp = 0;
    auto evaluator = std::make_shared<brm::RulesMctsRolloutEvaluator>(ctx, rp);
    brm::Mcts<brm::MvmctsState, brm::RulesMctsRolloutEvaluator> mcts(mp, evaluator, 0);
    const auto t0 = std::chrono::steady_clock::now();
    mcts.Search(state, iterations, n_roll);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const brm::JointAction best = mcts.ReturnBestAction();
    const int n_act = static_cast<int>(state->GetNumActions(0));
    const std::vector<double> st = mcts.GetRootStats(n_act);
    double visits = 0.0;
    for (int k = 0; k < n_act; ++k) {
      const double* s = &st[static_cast<size_t>(k) * (1 + ctx->V)];
      visits += s[0];
      std::printf("action %d visits %.0f mean %.6g\n", k, s[0], s[0] > 0 ? s[1] / s[0] : 0.0);
    }
    std::printf("best %d expected %d root visits %.0f iterations %d  %.3f s (%.1f us per iteration)\n",
                static_cast<int>(best[0]), expected, visits, mcts.NumIterations(), sec, 1e6 * sec / iterations);
    const bool ok = static_cast<int>(best[0]) == expected && static_cast<int>(visits) == iterations;
    std::printf("%s\n", ok ? "expectations hold" : "EXPECTATIONS FAIL");
    return ok ? 0 : 1;
  } catch (const brm::Error& err) {
    std::printf("error %d: %s\n", err.code, err.what());
    return 3;
  }
}
