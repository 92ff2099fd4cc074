// The reference's GridWorld threshold suite (test/grid_world_threshold_test.cpp:40-71) driven from
// C++: include/brm_mcts.hpp (host tree, leaf-parallel rounds) over include/brm_state.hpp (device
// states and rollout evaluator) over the C ABI.  TestRunner::RunTest(3000, 16) of
// gridworld/test_runner.cpp:17-52: search, execute the best joint action, ResetDepth, repeat.
//   mcts_test <rules.bin> <case: collision|pass|livelock> [batch_size] [rollouts_per_expansion]
// prints the final x positions; exit code 0 when the case's expectations hold.
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
  if (argc < 3) return 2;
  const std::vector<brm_rule> rules = read_file<brm_rule>(argv[1]);
  const std::string which = argv[2];
  const int batch = argc > 3 ? std::atoi(argv[3]) : 8;
  const int n_roll = argc > 4 ? std::atoi(argv[4]) : 16;
  const int kMergingPoint = 8;
  try {
    auto engine = std::make_shared<brm::Engine>(0);
    brm_gw_params p;
    std::memset(&p, 0, sizeof(p));
    // BaseTestEnv::MakeGridWorldStateParameters, gridworld/base_test_env.cpp:79-97
    p.n_agents = 3; p.state_x_length = 1000; p.ego_goal_reached_position = 1000; p.terminal_depth = 8;
    p.merging_point = kMergingPoint; p.reward_vec_size = 3; p.n_actions = 3;
    p.action_map[0] = 1; p.action_map[1] = 0; p.action_map[2] = -1;
    p.speed_deviation_weight = 5.0; p.acceleration_weight = 1.5; p.potential_weight = 0.0;
    auto ctx = std::make_shared<brm::GridWorldContext>(engine, p, rules);
    std::vector<uint8_t> q(ctx->A * ctx->R);
    for (int a = 0; a < ctx->A; ++a) {
      int s = 0;
      for (const brm_rule& r : rules) {
        bool specific = false;
        for (int k = 0; k < r.n_aps; ++k) specific |= r.ap_agent_specific[k] != 0;
        for (int j = 0; j < (specific ? ctx->A - 1 : 1); ++j) q[a * ctx->R + s++] = r.init_state;
      }
    }
    std::vector<int32_t> agents = {2, 1, 1, 1, 6, 1, 1, 1, 0, 1, 3, 3};  // base_test_env.cpp:29-43
    auto state = std::make_shared<brm::GridWorldState>(ctx, agents, 0, 0, q);
    // BaseTestEnv::MakeMctsParameters, gridworld/base_test_env.cpp:45-77
    brm::MctsParameters mp(3);
    mp.DISCOUNT_FACTOR = 0.95;
    mp.EXPLORATION_CONSTANT = {0.8, 0.8, 0.8};
    mp.LOWER_BOUND = {-1.0, -1.0, -100.0};
    mp.UPPER_BOUND = {0.0, 0.0, 0.0};
    mp.batch_size = batch;
    if (which == "collision") mp.THRESHOLD = {-1.0, -1.0, 1e300};
    else if (which == "pass") mp.THRESHOLD = {-0.37, -0.37, 1e300};
    else mp.THRESHOLD = {-0.1, -0.8, 1e300};
    brm_rollout_params rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.seed = 1000; rp.rollouts_per_leaf = n_roll; rp.max_steps = 30; rp.discount_factor = 0.95;
    rp.first_discount_exp = 0;
    auto evaluator = std::make_shared<brm::GridWorldRolloutEvaluator>(ctx, rp);
    typedef brm::Mcts<brm::GridWorldState, brm::GridWorldRolloutEvaluator> Search;
    const auto t0 = std::chrono::steady_clock::now();
    int steps = 0, iterations = 0;
    while (!state->IsTerminal() && steps < 16) {
      Search mcts(mp, evaluator, static_cast<uint64_t>(steps) << 32);  // a fresh tree per decision
      mcts.Search(state, 3000, n_roll);
      iterations += mcts.NumIterations();
      std::vector<brm::Reward> rewards;
      state = state->Execute(mcts.ReturnBestAction(), rewards);
      state->ResetDepth();
      ++steps;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const auto& x = state->GetAgentStates();
    const int x0 = x[0], x1 = x[4], x2 = x[8];
    std::printf("%s: steps %d iterations %d x %d %d %d  %.2f s (%.1f us per iteration)\n", which.c_str(), steps,
                iterations, x0, x1, x2, sec, 1e6 * sec / iterations);
    bool ok;
    if (which == "collision") ok = x0 == x2 && x1 > kMergingPoint;                                // :45-47
    else if (which == "pass") ok = x0 > kMergingPoint && x1 > kMergingPoint && x2 > kMergingPoint && x0 < x2 && x2 < x1;  // :55-60
    else ok = x0 < kMergingPoint && x1 > kMergingPoint && x2 < kMergingPoint;                      // :68-70
    std::printf("%s\n", ok ? "expectations hold" : "EXPECTATIONS FAIL");
    return ok ? 0 : 1;
  } catch (const brm::Error& err) {
    std::printf("error %d: %s\n", err.code, err.what());
    return 3;
  }
}
