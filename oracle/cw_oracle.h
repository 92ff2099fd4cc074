/* TEST INFRASTRUCTURE -- CPU oracle for the continuous-world (BARK-backed) hot path.
 * Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load oracle/libcw_oracle.so.
 *
 * PARITY UNPINNED for everything that lives in the un-vendored dependencies: the
 * reference's BARK-backed state MvmctsState (/root/reference/src/rules_mcts_state.cpp)
 * executes inside bark @ 9649bc11, rule-monitoring @ 187c125a and lexmamcts @ 24ab42ca
 * (/root/reference/util/deps.bzl:6-31), none of which is in /root/reference, and no
 * bazel/Eigen/boost/Spot exists here, so it cannot be compiled.  This file restates
 *   - term by term, what IS in the reference: MvmctsState::Execute (:57-128),
 *     GetNumActions (:129-141), GetTerminalReward (:148-164), EvaluateRules (:165-185),
 *     GetActionCost (:202-248), PotentialReward/Potential (:250-264), CheckTerminal
 *     (:266-285), parameter defaults (src/rules_mcts_state_parameters.cpp:36-78);
 *   - from the published algorithms, what is not: single-track kinematics with explicit
 *     Euler sub-steps (bark SingleTrackModel / BehaviorMPContinuousActions), rectangle
 *     collision (EvaluatorCollisionEgoAgent), polygon drivable area (EvaluatorDrivableArea),
 *     Frenet projection on lane centre lines, and the label functions safe-distance /
 *     preceding-agent / beyond-point / near / rightmost-lane.  These definitions are
 *     normative for this build and documented in DESIGN.md.
 * Pins that exist: the three scalar rewards of test/single_agent_test.cpp:143,152,159
 * (0 / -800 / -1000) and the action-count / frozen-position behaviour of
 * test/multi_agent_test.cpp:115-139, reproduced in tests/test_cw_oracle.py.
 *
 * Arithmetic runs in double, the reference's type (src/rules_mcts_state.cpp:204-264), on the
 * parameters as given (no rounding through float).  Every threshold comparison records how
 * close it came to flipping (`margin` outputs) -- a diagnostic, no test excludes anything on it.
 */
#ifndef ORACLE_CW_ORACLE_H_
#define ORACLE_CW_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CWO_MAX_AGENTS 16
#define CWO_MAX_LANES 8
#define CWO_MAX_PTS 64
#define CWO_MAX_ROAD 32
#define CWO_MAX_PRIMS 8
#define CWO_MAX_RULES 8
#define CWO_MAX_SLOTS 32
#define CWO_MAX_APS 4
#define CWO_MAX_STATES 16
#define CWO_VIOLATION 0xFF

/* label kinds; kinds >= 8 are per-other-agent (`name#0`) */
enum {
  CWO_LABEL_COLLISION_EGO = 0,
  CWO_LABEL_COLLISION_CORRIDOR = 1,
  CWO_LABEL_GOAL_REACHED = 2,
  CWO_LABEL_SAFE_DISTANCE_FRONT = 3,
  CWO_LABEL_MERGED_E = 4,
  CWO_LABEL_RIGHTMOST_LANE = 5,
  CWO_LABEL_ON_ROAD = 6,
  CWO_LABEL_MERGED_X = 8,
  CWO_LABEL_IN_DIRECT_FRONT_X = 9,
  CWO_LABEL_NEAR_X = 10
};

typedef struct {
  int32_t n_aps, n_states, init_state, priority, reset_on_violation;
  float weight;
  int32_t ap_label[CWO_MAX_APS];
  uint8_t trans[CWO_MAX_STATES * 16];
  uint8_t accepting[CWO_MAX_STATES];
} cwo_rule;

typedef struct {
  double front, rear, half_width; /* rectangle [-rear, front] x [-hw, hw] around the pose */
  int32_t has_goal, n_prims;
  double goal[6]; /* xmin, xmax, ymin, ymax, theta_lo, theta_hi */
  double prim_acc[CWO_MAX_PRIMS], prim_delta[CWO_MAX_PRIMS]; /* (a, steering angle delta) */
} cwo_agent;

typedef struct {
  int32_t n_pts, successor, flags, _pad; /* flags bit0: rightmost lane */
  double half_width, s_point;            /* beyond-point threshold along the lane */
  double pts[CWO_MAX_PTS][2];
} cwo_lane;

typedef struct {
  /* RulesMctsStateParameters (src/rules_mcts_state_parameters.hpp:33-46) */
  double collision_weight, out_of_map_weight, potential_weight, acceleration_weight,
      radial_acceleration_weight, desired_velocity_weight, lane_center_weight;
  double prediction_time_span, desired_velocity, discount_factor, goal_reward;
  int32_t reward_vector_size, horizon, use_rule_reward_for_ego_only;
  int32_t remove_agents_out_of_map; /* bark "World::remove_agents_out_of_map" (default 0) */
  /* dynamics (bark, un-vendored) */
  double wheel_base, integration_time_delta;
  /* label parameters */
  double sd_reaction_time, sd_max_decel_rear, sd_max_decel_front, near_distance;
  int32_t n_agents, n_mcts_agents, n_lanes, n_road, n_rules, _pad;
  cwo_agent agents[CWO_MAX_AGENTS];
  cwo_lane lanes[CWO_MAX_LANES];
  double road[CWO_MAX_ROAD][2];
  cwo_rule rules[CWO_MAX_RULES];
} cwo_config;

int32_t cwo_rules_per_agent(const cwo_config* cfg);

/* agent flag bits */
#define CWO_PRESENT 1
#define CWO_VALID 2

/* CheckTerminal of a state (src/rules_mcts_state.cpp:266-285). */
int32_t cwo_check_terminal(const cwo_config* cfg, const double* states /*[A][4] x,y,theta,v*/,
                           const uint8_t* flags /*[A]*/, int32_t horizon);

/* Replay T joint actions (per MCTS agent) from one state; stops once terminal.
 *  states_out [T][A][4] double, flags_out [T][A], terminal_out [T], q_out [T][M][R],
 *  viol_out [T][M][R], rewards_out [T][M][V], labels_out [T][M][4] (word 0 ego bits by
 *  label kind, words 1..3 masks over j for kinds 8,9,10), frenet_out [T][A][3]
 *  (lane, s, lat), margin_out [T] (smallest |lhs-rhs| over all threshold comparisons of
 *  the step), term_reward_out [M][V].  Returns the number of steps executed. */
int32_t cwo_replay(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                   int32_t horizon0, const uint8_t* q0, const uint8_t* actions, int32_t T,
                   double* states_out, uint8_t* flags_out, uint8_t* terminal_out, uint8_t* q_out,
                   uint32_t* viol_out, double* rewards_out, uint32_t* labels_out,
                   double* frenet_out, double* margin_out, double* term_reward_out);

/* Random rollouts with Philox actions (same draw rule as gw_oracle.h). */
int32_t cwo_rollout(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                    int32_t horizon0, const uint8_t* q0, uint64_t seed, uint64_t rollout_begin,
                    uint64_t n_rollouts, int32_t max_steps, double gamma,
                    int32_t first_discount_exp, double coop_factor, int32_t n_threads,
                    double* returns_sum /*[M][V], agents mixed with coop_factor*/,
                    uint64_t* viol_sum /*[M][R]*/, uint64_t* agent_steps);

/* value_i = own_i + coop_factor * sum_{j != i} own_j, in place on returns_sum [M][V]. */
void cwo_coop_mix(double* returns_sum, int32_t M, int32_t V, double coop_factor);

/* MvmctsState::EvaluateRules() (src/rules_mcts_state.cpp:186-197) on one state: q_out / viol_out
 * [M][R], rewards_out [M][V] (rule penalties only), labels_out [M][4]. */
int32_t cwo_evaluate_rules(const cwo_config* cfg, const double* states, const uint8_t* flags,
                           int32_t horizon, const uint8_t* q_in, uint8_t* q_out, uint32_t* viol_out,
                           double* rewards_out, uint32_t* labels_out);

/* Per-rollout outputs: returns [n][M][V], viol [n][M][R], q [n][M][R], states [n][A][4],
 * flags [n][A], steps [n], margin [n] (min over the rollout). */
int32_t cwo_rollout_detail(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                           int32_t horizon0, const uint8_t* q0, uint64_t seed,
                           uint64_t rollout_begin, uint64_t n_rollouts, int32_t max_steps,
                           double gamma, int32_t first_discount_exp, double* returns,
                           uint32_t* viol, uint8_t* q_final, double* states_final,
                           uint8_t* flags_final, int32_t* steps, double* margin);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_CW_ORACLE_H_ */
