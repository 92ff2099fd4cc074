/* brm.h -- C ABI of the B200 rollout engine for rule-aware multi-agent MCTS.
 *
 * This is the drop-in boundary for the simulation phase of
 * bark-simulator/planner-rules-mcts.  The reference has no FFI: its boundary is the
 * C++ CRTP plugin mvmcts::StateInterface<Impl> (Execute / GetNumActions / IsTerminal /
 * GetTerminalReward; /root/reference/src/rules_mcts_state.hpp:43-64,
 * /root/reference/gridworld/grid_world_state.hpp:28-71) called by
 * mvmcts::Mvmcts::Search and mvmcts::RandomHeuristic.  Each entry point below cites
 * the reference interface it replaces; include/brm_state.hpp is the thin C++ adapter
 * that re-exposes the StateInterface member set on top of this ABI, and
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions, no STL, no torch types.
 *   - every call returns BRM_OK (0) or a negative error; brm_last_error() gives text.
 *   - one brm_engine = one GPU + one CUDA stream; calls on one engine are not
 *     re-entrant (the reference is single-threaded too: a process-global
 *     std::mt19937, src/behavior_rules_mcts.hpp:154).
 *   - batches are structure-of-arrays: element [k][i] of an array documented as
 *     [K][n] lives at k*n + i, i = index of the state / leaf in the batch.
 *   - `mem` selects where the batch arrays live: BRM_MEM_HOST (copied in and out
 *     inside the call) or BRM_MEM_DEVICE (pointers valid on the engine's GPU; the
 *     call only enqueues work on the engine's stream and returns).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     BRM_ERR_NO_DEVICE when no sm_100-class GPU is usable.
 */
#ifndef BRM_H_
#define BRM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRM_ABI_VERSION 3

#define BRM_OK 0
#define BRM_ERR_INVALID (-1)
#define BRM_ERR_NO_DEVICE (-2)
#define BRM_ERR_CUDA (-3)
#define BRM_ERR_TERMINAL (-4) /* Execute on a terminal state (BARK_EXPECT_TRUE, src/rules_mcts_state.cpp:59) */
#define BRM_ERR_NOT_CONFIGURED (-5)

#define BRM_MEM_HOST 0
#define BRM_MEM_DEVICE 1

#define BRM_MAX_APS 4
#define BRM_MAX_DFA_STATES 16
#define BRM_MAX_RULES 8
#define BRM_MAX_REWARD 4
#define BRM_MAX_ACTIONS 8
#define BRM_VIOLATION 0xFF

typedef struct brm_engine brm_engine;

/* ---------------------------------------------------------------- engine ---- */

/* Create an engine on CUDA device `device`.  `cuda_stream` may be NULL (the engine
 * creates its own non-blocking stream) or an existing cudaStream_t to enqueue on
 * (e.g. the caller's torch stream). */
int brm_create(int device, void* cuda_stream, brm_engine** out);
void brm_destroy(brm_engine* e);
const char* brm_last_error(void);
int brm_abi_version(void);
/* sizeof of the ABI's structs as this library was compiled, in the order: brm_rule,
 * brm_rollout_params, brm_gw_params, brm_gw_states, brm_cw_params, brm_cw_agent, brm_cw_lane,
 * brm_cw_scenario, brm_cw_states, brm_cw_rollout_out, brm_cw_trace.  Fills up to n entries and
 * returns the number there are -- lets a foreign-language binding check its own layouts. */
int brm_abi_struct_sizes(int32_t* out, int32_t n);
/* Block until everything enqueued on the engine's stream has finished. */
int brm_synchronize(brm_engine* e);
/* Number of kernels this engine has launched since creation (bench `gpu_launches`). */
uint64_t brm_launch_count(const brm_engine* e);
/* Device-side timing of the rollout kernels on the engine's stream: elapsed ms and
 * launch count accumulated since the last reset (CUDA events around every rollout
 * kernel launch; requires brm_set_kernel_timing(e, 1)). */
int brm_set_kernel_timing(brm_engine* e, int enabled);
int brm_kernel_time(brm_engine* e, double* ms_total, uint64_t* launches, int reset);
/* Device properties the roofline needs. */
int brm_device_info(brm_engine* e, int* sm_count, int* sm_clock_khz, int* cc_major, int* cc_minor);

/* ----------------------------------------------------------- rule monitors --- */

/* A deterministic rule monitor = what ltl::RuleMonitor::MakeRule(formula, weight,
 * priority) yields (call sites gridworld/base_test_env.cpp:101-102,
 * test/single_agent_test.cpp:62), compiled on the host to a dense table
 * (planner_rules_mcts_b200.ltl.compile_rule).  Evaluate()/FinalTransit() contract as
 * used at src/rules_mcts_state.cpp:158-159,180-181 and
 * gridworld/grid_world_state.cpp:77-78,147-148:
 *   letter = sum_k value(ap_k) << (n_aps-1-k); next = trans[state*(1<<n_aps)+letter];
 *   next == BRM_VIOLATION: reward[priority] += weight, ++violations, state := init_state
 *   (on_violation 0 = reset) or unchanged (1 = stay); FinalTransit = accepting[state]
 *   ? 0 : weight.
 * ap_label[k] is a label kind of the world the rule is used in (BRM_GW_LABEL_* or
 * BRM_CW_LABEL_*); ap_agent_specific[k] != 0 marks a `label#0` placeholder: the rule
 * is then instantiated once per other agent (MakeRuleState(others, {}),
 * gridworld/base_test_env.cpp:129-136, src/behavior_rules_mcts.hpp:295-322). */
typedef struct {
  uint8_t n_aps;
  uint8_t n_states;
  uint8_t init_state;
  uint8_t priority;
  uint8_t on_violation; /* 0 reset to init_state, 1 stay */
  uint8_t _pad[3];
  float weight;
  uint8_t ap_label[BRM_MAX_APS];
  uint8_t ap_agent_specific[BRM_MAX_APS];
  uint8_t trans[BRM_MAX_DFA_STATES * (1 << BRM_MAX_APS)]; /* row stride 1<<n_aps */
  uint8_t accepting[BRM_MAX_DFA_STATES];
} brm_rule;

/* ---------------------------------------------- rollout policy (RandomHeuristic) */

/* mvmcts::RandomHeuristic settings (src/util.cpp:28-49, gridworld/base_test_env.cpp:50,73):
 * per rollout, uniform random joint actions until the state is terminal or
 * max_steps Executes were done; return = sum_k gamma^(k+first_discount_exp) r_k +
 * gamma^(K+first_discount_exp) GetTerminalReward().
 * Random actions: Philox4x32-10, key = seed, counter = (rollout_lo, rollout_hi,
 * step>>2, agent), word step&3, action = mulhi32(word, n_actions(agent)). */
typedef struct {
  uint64_t seed;
  uint64_t rollout_base; /* global id of rollout 0 of leaf 0; leaf l, rollout i has id base + l*rollouts_per_leaf + i */
  int32_t rollouts_per_leaf;
  int32_t max_steps;          /* MAX_NUMBER_OF_ITERATIONS, default 30 */
  double discount_factor;     /* DISCOUNT_FACTOR */
  int32_t first_discount_exp; /* 1: first step reward weighted by gamma (default), 0: by 1 */
  int32_t _pad;
  /* COOP_FACTOR (src/util.cpp:28-29; 0 in every test of the reference): returns_sum of agent i
   * is own_i + coop_factor * sum_{j != i} own_j.  Per-rollout outputs (ro_returns) stay unmixed. */
  double coop_factor;
} brm_rollout_params;

/* ------------------------------------------------------------- GridWorld ------ */

#define BRM_GW_MAX_AGENTS 8
#define BRM_GW_MAX_SLOTS 16 /* rule states per agent */

/* label kinds: the four evaluators of gridworld/base_test_env.cpp:104-117 */
#define BRM_GW_LABEL_COLLISION 0         /* EvaluatorLabelCollision   "collision"          */
#define BRM_GW_LABEL_MERGED_E 1          /* EvaluatorLabelEgoRange    "merged_e"           */
#define BRM_GW_LABEL_MERGED_X 2          /* EvaluatorLabelInRange     "merged_x#j"         */
#define BRM_GW_LABEL_IN_DIRECT_FRONT_X 3 /* EvaluatorLabelInDirectFront "in_direct_front_x#j" */

/* mirrors GridWorldStateParameter, gridworld/grid_world_state_parameter.h:15-29 */
typedef struct {
  int32_t n_agents; /* num_other_agents + 1 */
  int32_t state_x_length;
  int32_t ego_goal_reached_position;
  int32_t terminal_depth;
  int32_t merging_point;
  int32_t reward_vec_size;
  int32_t n_actions;
  int32_t action_map[BRM_MAX_ACTIONS];
  double speed_deviation_weight;
  double acceleration_weight;
  double potential_weight;
} brm_gw_params;

/* A batch of n GridWorld states (GridWorldState members, gridworld/grid_world_state.hpp:84-90).
 *   agents [A][n][4] int32 (x_pos, last_action, lane, init_lane)  -- one 16-byte vector per agent
 *   term   [n]       bit a = terminal_agents_[a]; bit 0 = IsTerminal()
 *   depth  [n]       depth_
 *   q      [A][R][n] automaton state of rule slot r of agent a (R = brm_gw_rules_per_agent)
 *   viol   [A][R][n] violation counters (may be NULL: not tracked)                        */
typedef struct {
  int32_t n;
  int32_t mem;
  int32_t* agents;
  uint8_t* term;
  int32_t* depth;
  uint8_t* q;
  uint32_t* viol;
} brm_gw_states;

/* Upload scenario constants and the common rules of every agent (rule-state layout
 * per agent as BaseTestEnv::GetAutomataVec, gridworld/base_test_env.cpp:122-140:
 * rules in order, agent-specific rules expanded per other agent ascending). */
int brm_gw_configure(brm_engine* e, const brm_gw_params* p, const brm_rule* rules, int32_t n_rules);
int32_t brm_gw_rules_per_agent(const brm_engine* e);

/* GridWorldState::Execute (gridworld/grid_world_state.cpp:45-90) for n states.
 *   actions [A][n] uint8 action indices; rewards [A][V][n] double.
 * Step rewards of agents that were already terminal are 0 (the reference leaves the
 * caller's slot untouched, :59-62).  `in` and `out` must not overlap (Execute is const and
 * returns a new state in the reference as well). */
int brm_gw_execute(brm_engine* e, const brm_gw_states* in, const uint8_t* actions,
                   brm_gw_states* out, double* rewards);
/* GridWorldState::GetNumActions (:154-159): out [A][n] */
int brm_gw_get_num_actions(brm_engine* e, const brm_gw_states* s, uint8_t* out);
/* GridWorldState::IsTerminal (:153): out [n] */
int brm_gw_is_terminal(brm_engine* e, const brm_gw_states* s, uint8_t* out);
/* GridWorldState::GetTerminalReward (:141-152): out [A][V][n] double */
int brm_gw_terminal_reward(brm_engine* e, const brm_gw_states* s, double* out);

/* The rollout loop of mvmcts::RandomHeuristic over GridWorldState (call site
 * gridworld/grid_world_env.h:40): rollouts_per_leaf rollouts from each
 * of leaves->n leaf states, fused on the device (K1 gridworld_rollout).
 *   returns_sum [n][A][V] double  sum over the leaf's rollouts of the discounted return
 *   viol_sum    [n][A][R] uint64  sum of final violation counters (NULL: skip)
 *   agent_steps [n]       uint64  executed agent-steps (non-terminal agents stepped)
 * Optional per-rollout outputs for parity checks (NULL: skip), rollout index
 * g = leaf*rollouts_per_leaf + i:
 *   ro_returns [N][A][V] double, ro_viol [N][A][R] uint8, ro_q [N][A][R] uint8,
 *   ro_agents [N][A][4] int32, ro_steps [N] int32.
 * Output arrays live where `leaves->mem` says. */
typedef struct {
  double* returns_sum;
  uint64_t* viol_sum;
  uint64_t* agent_steps;
  double* ro_returns;
  uint8_t* ro_viol;
  uint8_t* ro_q;
  int32_t* ro_agents;
  int32_t* ro_steps;
} brm_gw_rollout_out;

int brm_gw_rollout(brm_engine* e, const brm_gw_states* leaves, const brm_rollout_params* rp,
                   const brm_gw_rollout_out* out);

/* Deterministic replay of given action sequences (parity instrument; the closed loop
 * of gridworld/test_runner.cpp:28-42 without the search): B traces of T steps from
 * B initial states; a trace stops once its state is terminal.
 *   actions [B][T][A] uint8
 *   agents_out [B][T][A][4] int32, term_out [B][T] uint8 (bitmask), q_out [B][T][A][R] uint8,
 *   viol_out [B][T][A][R] uint32, rewards_out [B][T][A][V] double,
 *   labels_out [B][T][A][3] uint32 (word 0: bit0 collision, bit1 merged_e; word 1: merged_x
 *   mask over j; word 2: in_direct_front_x mask over j), term_reward_out [B][A][V] double,
 *   steps_out [B] int32.   All host or all device according to s0->mem. */
typedef struct {
  int32_t* agents_out;
  uint8_t* term_out;
  uint8_t* q_out;
  uint32_t* viol_out;
  double* rewards_out;
  uint32_t* labels_out;
  double* term_reward_out;
  int32_t* steps_out;
} brm_gw_trace;

int brm_gw_replay(brm_engine* e, const brm_gw_states* s0, const uint8_t* actions, int32_t T,
                  const brm_gw_trace* trace);

/* ------------------------------------------- continuous world (MvmctsState) ------ */

/* The BARK-backed state plugin MvmctsState (src/rules_mcts_state.{hpp,cpp}) on a
 * self-contained world model: agents are rectangles with single-track kinematics and
 * motion primitives (a, delta), the map is a set of lane centre polylines plus one road
 * polygon.  bark / ltl / mvmcts are un-vendored dependencies of the reference; DESIGN.md
 * states which definitions restate reference lines and which restate the published
 * algorithms ("parity unpinned"). */

#define BRM_CW_MAX_AGENTS 16
#define BRM_CW_MAX_LANES 8
#define BRM_CW_MAX_PTS 64    /* points per lane centre line */
#define BRM_CW_MAX_ROAD 32   /* road polygon vertices */
#define BRM_CW_MAX_PRIMS 8   /* motion primitives per agent */
#define BRM_CW_MAX_SLOTS 32  /* rule states per agent */

/* label kinds; kinds >= 8 are per-other-agent (`name#0`).  bark label functions named in
 * brackets (bark/world/evaluation/ltl/label_functions/, un-vendored). */
#define BRM_CW_LABEL_COLLISION_EGO 0       /* GenericEgoLabelFunction<EvaluatorCollisionEgoAgent>, test/single_agent_test.cpp:51-52 */
#define BRM_CW_LABEL_COLLISION_CORRIDOR 1  /* GenericEgoLabelFunction<EvaluatorDrivableArea>, :55-56 */
#define BRM_CW_LABEL_GOAL_REACHED 2        /* GenericEgoLabelFunction<EvaluatorGoalReached>, :53-54 */
#define BRM_CW_LABEL_SAFE_DISTANCE_FRONT 3 /* [SafeDistanceLabelFunction] true while the gap to the preceding agent is safe */
#define BRM_CW_LABEL_MERGED_E 4            /* [EgoBeyondPointLabelFunction] */
#define BRM_CW_LABEL_RIGHTMOST_LANE 5      /* [RightmostLaneLabelFunction] */
#define BRM_CW_LABEL_ON_ROAD 6             /* ego is on some lane */
#define BRM_CW_LABEL_MERGED_X 8            /* [AgentBeyondPointLabelFunction] `merged_x#0` */
#define BRM_CW_LABEL_IN_DIRECT_FRONT_X 9   /* [PrecedingAgentLabelFunction]   `in_direct_front_x#0` */
#define BRM_CW_LABEL_NEAR_X 10             /* [AgentNearLabelFunction]        `near_x#0` */

/* mirrors RulesMctsStateParameters (src/rules_mcts_state_parameters.hpp:33-46, defaults
 * src/rules_mcts_state_parameters.cpp:36-78) plus the dynamics / label constants that
 * live in bark's parameter server.  Doubles, like the reference's members: the device
 * computes the continuous world in FP64 (labels / automaton states / violation counts must
 * equal the reference's double arithmetic bit for bit). */
typedef struct {
  double collision_weight;            /* 0     */
  double out_of_map_weight;           /* -800  */
  double potential_weight;            /* 32    */
  double acceleration_weight;         /* 0     */
  double radial_acceleration_weight;  /* 0     */
  double desired_velocity_weight;     /* -5    */
  double lane_center_weight;          /* 0     */
  double prediction_time_span;        /* 0.3   */
  double desired_velocity;            /* 10    */
  double discount_factor;             /* 0.9 (used by PotentialReward, :250-257) */
  double goal_reward;                 /* 0     */
  int32_t reward_vector_size;         /* 1     */
  int32_t horizon;                    /* 30    */
  int32_t use_rule_reward_for_ego_only; /* false */
  int32_t remove_agents_out_of_map;   /* bark "World::remove_agents_out_of_map" (false): agents whose
                                         shape leaves the road polygon are removed by Predict; an MCTS
                                         agent that vanished earns OUT_OF_MAP_WEIGHT (:121-125) */
  double wheel_base;                  /* 2.7   SingleTrackModel */
  double integration_time_delta;      /* 0.02  BehaviorMotionPrimitives::IntegrationTimeDelta */
  double sd_reaction_time;            /* safe-distance label: reaction time delta        */
  double sd_max_decel_rear;           /* |a_r|                                           */
  double sd_max_decel_front;          /* |a_f|                                           */
  double near_distance;               /* near_x#j: centre distance below this            */
} brm_cw_params;

typedef struct {
  double front, rear, half_width; /* rectangle [-rear, front] x [-half_width, half_width] around the pose */
  double goal[6];                 /* xmin, xmax, ymin, ymax, theta_lo, theta_hi */
  int32_t has_goal;
  int32_t n_prims;                /* motion primitives; agents outside the MCTS agent list use primitive 0 */
  double prim_acc[BRM_CW_MAX_PRIMS];
  double prim_delta[BRM_CW_MAX_PRIMS];
} brm_cw_agent;

typedef struct {
  int32_t n_pts;
  int32_t successor; /* lane that continues this one (-1: none) */
  int32_t flags;     /* bit0: rightmost lane */
  int32_t _pad;
  double half_width;
  double s_point;    /* arc length of the merge point on this lane (beyond-point labels) */
  double pts[BRM_CW_MAX_PTS][2];
} brm_cw_lane;

/* One scenario.  Agents 0..n_mcts_agents-1 are the MCTS agents (ego = 0; the reference's
 * agent_idx list, src/behavior_rules_mcts.hpp:392-414); the rest are simulated only.
 * Coordinates must lie within +-BRM_CW_MAX_COORD metres (the fp32 culling boxes of the lane
 * projection are grown for the rounding of such coordinates). */
#define BRM_CW_MAX_COORD 4096.0
typedef struct {
  brm_cw_params params;
  int32_t n_agents, n_mcts_agents, n_lanes, n_road, n_rules, _pad;
  brm_cw_agent agents[BRM_CW_MAX_AGENTS];
  brm_cw_lane lanes[BRM_CW_MAX_LANES];
  double road[BRM_CW_MAX_ROAD][2];
  brm_rule rules[BRM_MAX_RULES]; /* common rules of every MCTS agent */
} brm_cw_scenario;

/* agent flag bits */
#define BRM_CW_PRESENT 1 /* agent exists in the world */
#define BRM_CW_VALID 2   /* ExecutionStatus::VALID (cleared after a collision, :116-120) */

/* A batch of n states of scenarios (MvmctsState members, src/rules_mcts_state.hpp:80-87).
 *   scenario [n]        index of the uploaded scenario each state lives in
 *   agents   [A][n][4]  double (x, y, theta, v): two 16-byte vectors per agent, consecutive
 *                       states consecutive in memory (bark State [t, x, y, theta, v] without t)
 *   flags    [A][n]     BRM_CW_PRESENT | BRM_CW_VALID
 *   horizon  [n]        remaining horizon_
 *   terminal [n]        is_terminal_state_ (CheckTerminal, :266-285; fill with brm_cw_check_terminal)
 *   q        [M][R][n]  automaton states;  viol [M][R][n] violation counters (may be NULL) */
typedef struct {
  int32_t n;
  int32_t mem;
  int32_t* scenario;
  double* agents;
  uint8_t* flags;
  int32_t* horizon;
  uint8_t* terminal;
  uint8_t* q;
  uint32_t* viol;
} brm_cw_states;

/* Upload n_scenarios scenarios (all with the same n_agents, n_mcts_agents,
 * reward_vector_size and rule list shapes; geometry, goals and primitives may differ). */
int brm_cw_configure(brm_engine* e, const brm_cw_scenario* scenarios, int32_t n_scenarios);
int32_t brm_cw_rules_per_agent(const brm_engine* e);

/* MvmctsState ctor's CheckTerminal (:52, :266-285): fills s->terminal. */
int brm_cw_check_terminal(brm_engine* e, brm_cw_states* s);
/* MvmctsState::Execute (:57-128): actions [M][n] uint8, rewards [M][V][n] double (the
 * reference's Reward is an Eigen::VectorXd).
 * BRM_MEM_HOST batches are checked first: BRM_ERR_TERMINAL if any state is terminal (:59). */
int brm_cw_execute(brm_engine* e, const brm_cw_states* in, const uint8_t* actions,
                   brm_cw_states* out, double* rewards);
/* MvmctsState::EvaluateRules() (:186-197), the call BehaviorRulesMcts::Plan makes on the root
 * state to carry the automata from the last planning step to the current world
 * (src/behavior_rules_mcts.hpp:211-214): labels of the states as they are, one transition per
 * rule state of every MCTS agent that is present and VALID.  `out` = `in` with the new automaton
 * states (and violation counters); rewards [M][V][n] double = the rule penalties only. */
int brm_cw_evaluate_rules(brm_engine* e, const brm_cw_states* in, brm_cw_states* out, double* rewards);
/* MvmctsState::GetNumActions (:129-141): out [M][n] */
int brm_cw_get_num_actions(brm_engine* e, const brm_cw_states* s, uint8_t* out);
/* MvmctsState::GetTerminalReward (:148-164): out [M][V][n] double */
int brm_cw_terminal_reward(brm_engine* e, const brm_cw_states* s, double* out);

/* mvmcts::RandomHeuristic over MvmctsState (instantiated at
 * src/behavior_rules_mcts.hpp:188-189), fused on the device (K3 rules_rollout):
 *   returns_sum [n][M][V] double, viol_sum [n][M][R] uint64, agent_steps [n] uint64
 *   (agents integrated: present and VALID, MCTS or not).
 * Optional per-rollout outputs: ro_returns [N][M][V] double, ro_viol [N][M][R] uint8,
 * ro_q [N][M][R] uint8, ro_agents [N][A][4] double, ro_flags [N][A] uint8, ro_steps [N] int32. */
typedef struct {
  double* returns_sum;
  uint64_t* viol_sum;
  uint64_t* agent_steps;
  double* ro_returns;
  uint8_t* ro_viol;
  uint8_t* ro_q;
  double* ro_agents;
  uint8_t* ro_flags;
  int32_t* ro_steps;
} brm_cw_rollout_out;

int brm_cw_rollout(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                   const brm_cw_rollout_out* out);

/* Replay of given action sequences with a full trace (K6): actions [B][T][M];
 *   agents_out [B][T][A][4] double, flags_out [B][T][A], terminal_out [B][T],
 *   q_out [B][T][M][R], viol_out [B][T][M][R] uint32, rewards_out [B][T][M][V] double,
 *   labels_out [B][T][M][4] uint32 (word 0: ego bits by label kind; words 1..3: masks over j
 *   for kinds 8, 9, 10), frenet_out [B][T][A][3] double (lane, s, lat),
 *   term_reward_out [B][M][V] double, steps_out [B]. */
typedef struct {
  double* agents_out;
  uint8_t* flags_out;
  uint8_t* terminal_out;
  uint8_t* q_out;
  uint32_t* viol_out;
  double* rewards_out;
  uint32_t* labels_out;
  double* frenet_out;
  double* term_reward_out;
  int32_t* steps_out;
} brm_cw_trace;

int brm_cw_replay(brm_engine* e, const brm_cw_states* s0, const uint8_t* actions, int32_t T,
                  const brm_cw_trace* trace);

/* ---------------------------------------------------- root statistics merge ---- */

/* Root-parallel merge (new: the reference has a single tree; what must be summed is
 * the root's per-agent, per-action visit count and value-vector sum read by
 * ReturnBestAction / GetEgoIntNode().GetExpectedRewards(), gridworld/grid_world_env.h:33,38).
 * Fold per-leaf rollout sums into root statistics on the device (K5 root_stats_reduce):
 *   leaf_first_action [n][A] uint8 : the root joint action each leaf descends from
 *   returns_sum [n][A][V] double, rollouts_per_leaf
 *   stats (device or host per `mem`) [A][n_actions][1+V] double : (visits, value sums),
 *   ACCUMULATED into (zero it first for a fresh decision).
 * The packed `stats` buffer is what ranks allreduce(sum) over NCCL. */
int brm_root_stats_reduce(brm_engine* e, int32_t mem, int32_t n_leaves, int32_t n_agents,
                          int32_t n_actions, int32_t reward_vec_size,
                          const uint8_t* leaf_first_action, const double* returns_sum,
                          int32_t rollouts_per_leaf, double* stats);

/* Root-parallel merge across GPUs (one engine = one GPU = one rank).  The
 * communicator is NCCL (bound at run time; NVLink / NVSwitch inside a box).  Rank 0 makes an
 * id with brm_comm_unique_id and hands it to the other ranks by any side channel (MPI, a file,
 * torch.distributed ...); every rank then calls brm_comm_init with the same id -- a collective
 * call, like ncclCommInitRank.  Several engines of one process (one per GPU) must call it from
 * different threads.  With a communicator, brm_allreduce_root_stats sums `count` doubles over
 * the ranks in place on the engine's stream, and the brm_*_rollout_root_stats calls below do
 * that merge themselves before anything is copied back: every rank receives the statistics
 * of the whole box.  brm_comm_size: ranks of the engine's communicator (1 without one). */
#define BRM_COMM_ID_BYTES 128
int brm_comm_unique_id(uint8_t id[BRM_COMM_ID_BYTES]);
int brm_comm_init(brm_engine* e, const uint8_t id[BRM_COMM_ID_BYTES], int32_t rank, int32_t n_ranks);
int brm_comm_destroy(brm_engine* e);
int brm_comm_size(const brm_engine* e);
int brm_allreduce_root_stats(brm_engine* e, int32_t mem, double* stats, int64_t count);

/* One decision of a root-parallel search in ONE call: brm_*_rollout followed by
 * brm_root_stats_reduce on the device, before anything is copied back (host batches: one upload,
 * three launches, one download, one synchronisation).  out->returns_sum may be NULL for host
 * batches when only the statistics are wanted.  stats is ACCUMULATED into, like
 * brm_root_stats_reduce; leaf_first_action [n][A] uint8 (A = MCTS agents), every entry
 * < n_actions; n_actions must be at least the number of actions (motion primitives) an agent
 * has in the configured world (BRM_ERR_INVALID otherwise).  With a communicator (brm_comm_init)
 * the statistics the call returns are the SUM over all ranks of (the `stats` passed in + this
 * decision's rollouts): pass zeros for a fresh decision. */
int brm_gw_rollout_root_stats(brm_engine* e, const brm_gw_states* leaves, const brm_rollout_params* rp,
                              const brm_gw_rollout_out* out, const uint8_t* leaf_first_action,
                              int32_t n_actions, double* stats);
int brm_cw_rollout_root_stats(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                              const brm_cw_rollout_out* out, const uint8_t* leaf_first_action,
                              int32_t n_actions, double* stats);


#ifdef __cplusplus
}
#endif

#endif /* BRM_H_ */
