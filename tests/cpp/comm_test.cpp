// Root-parallel merge through the C ABI from C++: one engine per GPU in ONE process (a thread
// each), brm_comm_init over an id made by brm_comm_unique_id, then a decision per rank with
// brm_gw_rollout_root_stats -- every rank must receive the sum of all ranks' statistics.
//   comm_test <rules.bin> <n_gpus> <rollouts_per_leaf>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>

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
  if (argc < 4) return 2;
  const std::vector<brm_rule> rules = read_file<brm_rule>(argv[1]);
  const int n = std::atoi(argv[2]), n_rollouts = std::atoi(argv[3]);
  brm_gw_params p;
  std::memset(&p, 0, sizeof(p));
  p.n_agents = 3; p.state_x_length = 1000; p.ego_goal_reached_position = 1000; p.terminal_depth = 8;
  p.merging_point = 8; p.reward_vec_size = 3; p.n_actions = 3;
  p.action_map[0] = 1; p.action_map[1] = 0; p.action_map[2] = -1;
  p.speed_deviation_weight = 5.0; p.acceleration_weight = 1.5; p.potential_weight = 0.0;
  uint8_t id[BRM_COMM_ID_BYTES];
  brm::check(brm_comm_unique_id(id));
  const size_t n_stats = 3 * 3 * (1 + 3);
  std::vector<std::vector<double>> own(n, std::vector<double>(n_stats, 0.0)), merged(n, std::vector<double>(n_stats, 0.0));
  std::vector<int> rc(n, 0);
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i)
    th.emplace_back([&, i]() {
      try {
        auto engine = std::make_shared<brm::Engine>(i);
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
        std::vector<int32_t> agents = {2, 1, 1, 1, 6, 1, 1, 1, 0, 1, 3, 3};
        auto root = std::make_shared<brm::GridWorldState>(ctx, agents, 0, 0, q);
        brm_rollout_params rp;
        std::memset(&rp, 0, sizeof(rp));
        rp.seed = 1000; rp.rollouts_per_leaf = n_rollouts; rp.max_steps = 30; rp.discount_factor = 0.95;
        rp.first_discount_exp = 1;
        brm::GridWorldRolloutEvaluator ev(ctx, rp);
        std::vector<uint8_t> first = {0, 1, 2, 2, 1, 0};
        const uint64_t base = 1000000ull * i;  // every rank its own rollouts
        ev.Evaluate({root, root}, base, &first, 3, &own[i]);  // no communicator yet: this rank's statistics
        brm::check(brm_comm_init(engine->get(), id, i, n));
        if (brm_comm_size(engine->get()) != n) throw brm::Error(-1, "brm_comm_size");
        ev.Evaluate({root, root}, base, &first, 3, &merged[i]);  // the same decision, merged across the ranks
        brm::check(brm_comm_destroy(engine->get()));
      } catch (const brm::Error& err) {
        std::printf("rank %d: error %d: %s\n", i, err.code, err.what());
        rc[i] = 1;
      }
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i]) return 1;
  int bad = 0;
  for (size_t k = 0; k < n_stats; ++k) {
    double want = 0.0;
    for (int i = 0; i < n; ++i) want += own[i][k];
    for (int i = 0; i < n; ++i)
      if (std::fabs(merged[i][k] - want) > 1e-9 * std::fmax(1.0, std::fabs(want))) ++bad;
  }
  double visits = 0.0;
  for (int a = 0; a < 3; ++a)
    for (int k = 0; k < 3; ++k) visits += merged[0][(a * 3 + k) * 4];
  std::printf("ranks %d mismatches %d visits %.0f differ %d\n", n, bad, visits,
              n > 1 && own[0] != own[n - 1] ? 1 : 0);
  return bad ? 1 : 0;
}
