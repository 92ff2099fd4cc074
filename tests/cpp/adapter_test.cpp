// Exercises include/brm_state.hpp (the StateInterface adapter over the C ABI) from C++.
//   adapter_test nogpu                 -> engine creation must fail loudly (no CPU fallback)
//   adapter_test gw <rules.bin> <n>    -> GridWorld closed loop + batched rollout evaluation
//   adapter_test cw <scenario.bin>     -> MvmctsState pins of test/single_agent_test.cpp
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "brm_state.hpp"

template <class T>
static std::vector<T> read_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<T> out(raw.size() / sizeof(T));
  std::memcpy(out.data(), raw.data(), out.size() * sizeof(T));
  return out;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "nogpu";
  if (mode == "nogpu") {
    try {
      brm::Engine e(0);
    } catch (const brm::Error& err) {
      std::printf("error code %d: %s\n", err.code, err.what());
      return err.code == BRM_ERR_NO_DEVICE ? 0 : 2;
    }
    std::printf("engine created (a GPU is present)\n");
    return 0;
  }
  auto engine = std::make_shared<brm::Engine>(0);
  if (mode == "gw") {
    std::vector<brm_rule> rules = read_file<brm_rule>(argv[2]);
    const int n_rollouts = std::atoi(argv[3]);
    brm_gw_params p;
    std::memset(&p, 0, sizeof(p));
    // BaseTestEnv::MakeGridWorldStateParameters, gridworld/base_test_env.cpp:79-97
    p.n_agents = 3; p.state_x_length = 1000; p.ego_goal_reached_position = 1000; p.terminal_depth = 8;
    p.merging_point = 8; p.reward_vec_size = 3; p.n_actions = 3;
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
    // BaseTestEnv initial state (base_test_env.cpp:29-43)
    std::vector<int32_t> agents = {2, 1, 1, 1, 6, 1, 1, 1, 0, 1, 3, 3};
    auto s = std::make_shared<brm::GridWorldState>(ctx, agents, 0, 0, q);
    std::vector<brm::Reward> rewards;
    int step = 0;
    // the closed loop of gridworld/test_runner.cpp:28-42 with a fixed plan: ego waits twice
    while (!s->IsTerminal() && step < 20) {
      brm::JointAction ja = {step < 2 ? 1u : 0u, 0u, 0u};
      s = s->Execute(ja, rewards);
      s->ResetDepth();
      std::printf("step %d x %d %d %d r0 %.17g %.17g %.17g\n", step, s->GetAgentStates()[0], s->GetAgentStates()[4],
                  s->GetAgentStates()[8], rewards[0][0], rewards[0][1], rewards[0][2]);
      ++step;
    }
    std::printf("terminal %d steps %d num_actions %zu %s", s->IsTerminal() ? 1 : 0, step, s->GetNumActions(0),
                s->PrintState().c_str());
    auto tr = s->GetTerminalReward();
    std::printf("terminal_reward %.17g %.17g %.17g\n", tr[0][0], tr[0][1], tr[0][2]);
    // batched rollouts from the initial state (RandomHeuristic replacement)
    brm_rollout_params rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.seed = 1000; rp.rollouts_per_leaf = n_rollouts; rp.max_steps = 30; rp.discount_factor = 0.95;
    rp.first_discount_exp = 1;
    brm::GridWorldRolloutEvaluator ev(ctx, rp);
    auto root = std::make_shared<brm::GridWorldState>(ctx, agents, 0, 0, q);
    auto means = ev.Evaluate({root, root}, 0);
    for (size_t l = 0; l < means.size(); ++l)
      for (int a = 0; a < ctx->A; ++a)
        std::printf("leaf %zu agent %d mean %.17g %.17g %.17g\n", l, a, means[l][a][0], means[l][a][1], means[l][a][2]);
    // the same rollouts with the root statistics folded in the same call
    std::vector<uint8_t> first = {0, 1, 2, 2, 1, 0};  // [leaf][agent]
    std::vector<double> stats(static_cast<size_t>(ctx->A) * 3 * (1 + ctx->V), 0.0);
    ev.Evaluate({root, root}, 0, &first, 3, &stats);
    for (int a = 0; a < ctx->A; ++a)
      for (int k = 0; k < 3; ++k) {
        const double* st = &stats[(static_cast<size_t>(a) * 3 + k) * (1 + ctx->V)];
        std::printf("stats agent %d action %d visits %.17g sums %.17g %.17g %.17g\n", a, k, st[0], st[1], st[2], st[3]);
      }
    return 0;
  }
  if (mode == "cw") {
    std::vector<brm_cw_scenario> scn = read_file<brm_cw_scenario>(argv[2]);
    auto ctx = std::make_shared<brm::RulesMctsContext>(engine, scn);
    std::vector<uint8_t> q(ctx->M * ctx->R, 0), flags(ctx->A, BRM_CW_PRESENT | BRM_CW_VALID);
    std::vector<double> agents = {0.0, -1.75, 0.0, 5.0, 11.0, -1.75, 0.0, 5.0};
    brm::MvmctsState s0(ctx, 0, agents, flags, 20, q);
    std::vector<brm::Reward> rewards;
    std::printf("terminal0 %d num_actions %zu\n", s0.IsTerminal() ? 1 : 0, s0.GetNumActions(0));
    auto s1 = s0.Execute({0}, rewards);   // test/single_agent_test.cpp:136-143
    double total = rewards[0][0] + s1->GetTerminalReward()[0][0];
    std::printf("a0 terminal %d total %.9g\n", s1->IsTerminal() ? 1 : 0, total);
    auto s2 = s1->Execute({1}, rewards);  // :146-152
    total += rewards[0][0] + s2->GetTerminalReward()[0][0];
    std::printf("a1 terminal %d total %.9g\n", s2->IsTerminal() ? 1 : 0, total);
    auto s3 = s0.Execute({2}, rewards);   // :154-159
    std::printf("a2 terminal %d total %.9g num_actions %zu\n", s3->IsTerminal() ? 1 : 0,
                rewards[0][0] + s3->GetTerminalReward()[0][0], s3->GetNumActions(0));
    try {
      s3->Execute({0}, rewards);
      std::printf("execute on terminal: no error\n");
    } catch (const brm::Error& err) {
      std::printf("execute on terminal: error %d\n", err.code);
    }
    // EvaluateRules() on the collided state: the `G !collision_ego` monitor fires once more
    brm::MvmctsState s4(*s3);
    auto rule_rewards = s4.EvaluateRules();
    std::printf("evaluate_rules on collided state %.9g, on the initial state %.9g\n", rule_rewards[0][0],
                brm::MvmctsState(s0).EvaluateRules()[0][0]);
    return 0;
  }
  return 1;
}
