/* TEST INFRASTRUCTURE -- CPU oracle for the GridWorld hot path.  Not part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the libraries built from this directory.
 *
 * Two implementations share this C interface:
 *   gwref_*  oracle/_ref/libgw_ref.so  = the reference's OWN unmodified sources
 *            (/root/reference/gridworld/grid_world_state.cpp, common.cpp,
 *            label_evaluator/ sources) compiled against oracle/shim/ headers and driven
 *            by oracle/ref_gridworld_driver.cpp.  Step, label and action-cost
 *            arithmetic are the reference's code; the rule automaton is the
 *            table-driven monitor of oracle/shim/ltl/rule_monitor.h.
 *   gwo_*    oracle/libgw_oracle.so    = plain-C restatement (oracle/gw_oracle.c),
 *            each function citing the reference file:line it follows.  Pinned
 *            against gwref_* on random traces and against the 23 known answers of
 *            /root/reference/gridworld/label_evaluator/test/label_test.cpp.
 */
#ifndef ORACLE_GW_ORACLE_H_
#define ORACLE_GW_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWO_MAX_AGENTS 8
#define GWO_MAX_APS 4
#define GWO_MAX_STATES 16
#define GWO_MAX_RULES 8
#define GWO_MAX_ACTIONS 8
#define GWO_VIOLATION 0xFF

/* label kinds (the 4 evaluators of gridworld/base_test_env.cpp:104-117) */
enum {
  GWO_LABEL_COLLISION = 0,         /* ego label  "collision"            */
  GWO_LABEL_MERGED_E = 1,          /* ego label  "merged_e"             */
  GWO_LABEL_MERGED_X = 2,          /* per-other  "merged_x#j"           */
  GWO_LABEL_IN_DIRECT_FRONT_X = 3  /* per-other  "in_direct_front_x#j"  */
};

typedef struct {
  int32_t n_aps;
  int32_t n_states;
  int32_t init_state;
  int32_t priority;
  int32_t reset_on_violation;
  float weight;
  int32_t ap_label[GWO_MAX_APS];            /* label kind of AP k; kinds >= 2 bind `#0` */
  uint8_t trans[GWO_MAX_STATES * 16];       /* [state][letter] -> next | GWO_VIOLATION; row stride 1<<n_aps */
  uint8_t accepting[GWO_MAX_STATES];        /* 1: FinalTransit = 0, 0: FinalTransit = weight */
} gwo_rule;

typedef struct {
  int32_t n_agents;
  int32_t state_x_length;
  int32_t ego_goal_reached_position;
  int32_t terminal_depth;
  int32_t merging_point;
  int32_t reward_vec_size;
  int32_t n_actions;
  int32_t action_map[GWO_MAX_ACTIONS];
  double speed_deviation_weight;
  double acceleration_weight;
  double potential_weight;
  int32_t n_rules;
  int32_t _pad;
  gwo_rule rules[GWO_MAX_RULES]; /* common rules of every agent, in Rule-enum order */
} gwo_config;

/* rule states per agent: sum over rules of (agent-specific ? n_agents-1 : 1) */
int32_t gwo_rules_per_agent(const gwo_config* cfg);
int32_t gwref_rules_per_agent(const gwo_config* cfg);

/* Label bitmask layout per perspective agent i (own bits are 0):
 *   bit 0 collision, bit 1 merged_e, bit 2+j merged_x#j, bit 2+A+j in_direct_front_x#j */

/* Replay T joint actions from one initial state; stops early once the state is
 * terminal (ego terminal).  All outputs are indexed [step][agent]...; the step
 * reward of an already-terminal agent is defined as 0 (the reference leaves the
 * slot untouched: gridworld/grid_world_state.cpp:59-62).
 *   agents0      [A][4]  (x_pos, last_action, lane, init_lane)
 *   term0        [A]
 *   q0           [A][R] or NULL (initial automaton states)
 *   actions      [T][A]  action indices
 *   agents_out   [T][A][4]
 *   term_out     [T][A]
 *   q_out        [T][A][R]
 *   viol_out     [T][A][R] cumulative violation counters
 *   rewards_out  [T][A][V]
 *   labels_out   [T][A]   label bitmask of the NEW state seen from agent a
 *   term_reward_out [A][V] GetTerminalReward() of the last state reached
 * returns number of steps executed.                                              */
#define GWO_REPLAY_ARGS                                                              \
  const gwo_config *cfg, const int32_t *agents0, const uint8_t *term0, int32_t depth0, \
      const uint8_t *q0, const uint8_t *actions, int32_t T, int32_t *agents_out,     \
      uint8_t *term_out, uint8_t *q_out, uint32_t *viol_out, double *rewards_out,    \
      uint32_t *labels_out, double *term_reward_out

int32_t gwo_replay(GWO_REPLAY_ARGS);
int32_t gwref_replay(GWO_REPLAY_ARGS);

/* Random rollouts (the mvmcts::RandomHeuristic loop) from
 * one leaf state with Philox4x32-10 actions:
 *   action(rollout r, step k, agent a) = mulhi32(philox(key=seed, ctr=(r_lo, r_hi, k>>2, a))[k&3], n_actions(a))
 * Return per rollout: sum_k gamma^(k+e) r_k + gamma^(K+e) terminal_reward, e = first_discount_exp.
 *   returns_sum   [A][V]  sum over rollouts (double)
 *   viol_sum      [A][R]  sum over rollouts of final violation counters
 *   agent_steps   executed agent-steps (non-terminal agents stepped)
 * n_threads > 1 splits the rollout range over std::thread/pthreads workers.        */
#define GWO_ROLLOUT_ARGS                                                             \
  const gwo_config *cfg, const int32_t *agents0, const uint8_t *term0, int32_t depth0, \
      const uint8_t *q0, uint64_t seed, uint64_t rollout_begin, uint64_t n_rollouts, \
      int32_t max_steps, double gamma, int32_t first_discount_exp, int32_t n_threads, \
      double *returns_sum, uint64_t *viol_sum, uint64_t *agent_steps

int32_t gwo_rollout(GWO_ROLLOUT_ARGS);
int32_t gwref_rollout(GWO_ROLLOUT_ARGS);

/* Per-rollout outputs for exact comparison (single thread):
 *   returns [n][A][V] doubles, viol [n][A][R], final q [n][A][R], final agents [n][A][4], steps [n] */
int32_t gwo_rollout_detail(const gwo_config* cfg, const int32_t* agents0, const uint8_t* term0,
                           int32_t depth0, const uint8_t* q0, uint64_t seed,
                           uint64_t rollout_begin, uint64_t n_rollouts, int32_t max_steps,
                           double gamma, int32_t first_discount_exp, double* returns,
                           uint32_t* viol, uint8_t* q_final, int32_t* agents_final,
                           int32_t* steps);

/* Philox4x32-10 known-answer access for tests */
void gwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* label unit checks used against label_test.cpp's known answers */
int32_t gwo_at_position(int32_t x, int32_t last_action, int32_t position);
int32_t gwo_check_collision(int32_t ax, int32_t alast, int32_t alane, int32_t bx, int32_t blast,
                            int32_t blane);
int32_t gwo_label_collision(int32_t merging_point, int32_t ex, int32_t elast, int32_t elane,
                            int32_t ox, int32_t olast, int32_t olane);
int32_t gwref_at_position(int32_t x, int32_t last_action, int32_t position);
int32_t gwref_label_collision(int32_t merging_point, int32_t ex, int32_t elast, int32_t elane,
                              int32_t ox, int32_t olast, int32_t olane);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_GW_ORACLE_H_ */
