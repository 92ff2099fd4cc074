/* TEST INFRASTRUCTURE -- plain-C restatement of the reference's GridWorld hot path.
 * Not part of the product (see oracle/gw_oracle.h).  Every function cites the
 * reference file:line (relative to /root/reference) it restates.  Pinned by
 * tests/test_gw_oracle.py against oracle/_ref/libgw_ref.so (the reference's own
 * code) on random traces and against label_test.cpp's 23 known answers. */
#include "oracle/gw_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle/philox.h"

typedef struct {
  int32_t x[GWO_MAX_AGENTS], last[GWO_MAX_AGENTS], lane[GWO_MAX_AGENTS], init_lane[GWO_MAX_AGENTS];
  uint8_t term[GWO_MAX_AGENTS];
  int32_t depth;
  uint8_t q[GWO_MAX_AGENTS][32];
  uint32_t viol[GWO_MAX_AGENTS][32];
} gw_state;

typedef struct {
  int32_t rule;  /* index into cfg->rules */
  int32_t other; /* bound other agent, -1 for ego-only rules */
} gw_slot;

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static int rule_is_specific(const gwo_rule* r) {
  for (int a = 0; a < r->n_aps; ++a)
    if (r->ap_label[a] >= 2) return 1;
  return 0;
}

/* rule-state layout per agent: rules in Rule-enum order, agent-specific rules expand
 * to one instance per other agent in ascending index order
 * (gridworld/base_test_env.cpp:122-140, multimap<Rule,RuleState> iteration order). */
static int build_slots(const gwo_config* cfg, int agent, gw_slot* slots) {
  int n = 0;
  for (int k = 0; k < cfg->n_rules; ++k) {
    if (rule_is_specific(&cfg->rules[k])) {
      for (int j = 0; j < cfg->n_agents; ++j)
        if (j != agent) {
          slots[n].rule = k;
          slots[n].other = j;
          ++n;
        }
    } else {
      slots[n].rule = k;
      slots[n].other = -1;
      ++n;
    }
  }
  return n;
}

int32_t gwo_rules_per_agent(const gwo_config* cfg) {
  gw_slot s[32];
  return build_slots(cfg, 0, s);
}

/* gridworld/label_evaluator/evaluator_label_at_position.cpp:16-28 */
int32_t gwo_at_position(int32_t x, int32_t last_action, int32_t p) {
  const int prev = x - last_action;
  return (prev < p && x >= p) || (prev > p && x <= p) || x == p;
}

/* gridworld/label_evaluator/evaluator_label_collision.cpp:14-31 */
int32_t gwo_check_collision(int32_t ax, int32_t alast, int32_t alane, int32_t bx, int32_t blast,
                            int32_t blane) {
  const int o_prev = bx - blast;
  const int upper_o = imax(o_prev, bx), lower_o = imin(o_prev, bx);
  const int e_prev = ax - alast;
  const int upper_e = imax(e_prev, ax), lower_e = imin(e_prev, ax);
  const int overlap = imin(upper_e, upper_o) - imax(lower_e, lower_o);
  int is_overlapping = 0;
  if (!((blast != 0 && bx == e_prev) || (alast != 0 && ax == o_prev)) || overlap != 0)
    is_overlapping = overlap >= 0;
  return (blane == alane) && (is_overlapping || ax == bx);
}

/* gridworld/label_evaluator/evaluator_label_collision.cpp:36-54, two-agent form */
int32_t gwo_label_collision(int32_t mp, int32_t ex, int32_t elast, int32_t elane, int32_t ox,
                            int32_t olast, int32_t olane) {
  return (gwo_at_position(ex, elast, mp) && gwo_at_position(ox, olast, mp)) ||
         gwo_check_collision(ex, elast, elane, ox, olast, olane);
}

/* label bitmask of state s from the perspective of agent i
 * (collision: evaluator_label_collision.cpp:36-54; merged_e: evaluator_label_ego_range.cpp:13-17;
 *  merged_x: evaluator_label_in_range.cpp:17-21; in_direct_front_x:
 *  evaluator_label_in_direct_front.cpp:15-34; per-other expansion:
 *  evaluator_label_multi_agent.cpp:15-23) */
static uint32_t labels_of(const gwo_config* cfg, const gw_state* s, int i) {
  const int A = cfg->n_agents, mp = cfg->merging_point, xl = cfg->state_x_length;
  uint32_t bits = 0;
  const int ego_at_xing = gwo_at_position(s->x[i], s->last[i], mp);
  for (int j = 0; j < A; ++j) {
    if (j == i) continue;
    const int other_at_xing = gwo_at_position(s->x[j], s->last[j], mp);
    if ((ego_at_xing && other_at_xing) ||
        gwo_check_collision(s->x[i], s->last[i], s->lane[i], s->x[j], s->last[j], s->lane[j])) {
      bits |= 1u;
      break;
    }
  }
  if (s->x[i] >= mp && s->x[i] <= xl) bits |= 2u;
  for (int j = 0; j < A; ++j) {
    if (j == i) continue;
    if (s->x[j] >= mp && s->x[j] <= xl) bits |= 1u << (2 + j);
    int front = (s->lane[j] == s->lane[i]) && (s->x[j] > s->x[i]);
    if (front) {
      for (int k = 0; k < A; ++k) {
        if (k == i || k == j) continue;
        if (s->x[k] >= s->x[i] && s->x[k] <= s->x[j] && s->lane[k] == s->lane[i]) {
          front = 0;
          break;
        }
      }
    }
    if (front) bits |= 1u << (2 + A + j);
  }
  return bits;
}

static int label_value(const gwo_config* cfg, uint32_t bits, int kind, int other) {
  switch (kind) {
    case GWO_LABEL_COLLISION: return bits & 1u;
    case GWO_LABEL_MERGED_E: return (bits >> 1) & 1u;
    case GWO_LABEL_MERGED_X: return (bits >> (2 + other)) & 1u;
    default: return (bits >> (2 + cfg->n_agents + other)) & 1u;
  }
}

/* one automaton step: contract of ltl::RuleMonitor::Evaluate as used at
 * gridworld/grid_world_state.cpp:76-79 (semantics: SURVEY.md Appendix B) */
static double rule_step(const gwo_config* cfg, const gw_slot* slot, uint32_t bits, uint8_t* q,
                        uint32_t* viol) {
  const gwo_rule* r = &cfg->rules[slot->rule];
  int letter = 0;
  for (int k = 0; k < r->n_aps; ++k)
    letter |= (label_value(cfg, bits, r->ap_label[k], slot->other) ? 1 : 0) << (r->n_aps - 1 - k);
  const uint8_t next = r->trans[(*q) * (1 << r->n_aps) + letter];
  if (next == GWO_VIOLATION) {
    ++*viol;
    if (r->reset_on_violation) *q = (uint8_t)r->init_state;
    return (double)r->weight;
  }
  *q = next;
  return 0.0;
}

/* GridWorldState::Execute, gridworld/grid_world_state.cpp:45-90 (+ Step :91-119,
 * SetCollisionPositions :211-226, GetActionCost :120-132) */
static void gw_execute(const gwo_config* cfg, const gw_state* in, const uint8_t* actions,
                       gw_state* out, double* rewards, uint32_t* labels) {
  const int A = cfg->n_agents, V = cfg->reward_vec_size;
  *out = *in;
  /* Step (:91-119) */
  for (int i = 0; i < A; ++i) {
    if (in->term[i]) continue;
    const int d = cfg->action_map[actions[i]];
    const int nx = in->x[i] + d;
    out->x[i] = nx;
    out->lane[i] = (nx >= cfg->merging_point) ? 0 : in->init_lane[i];
    out->last[i] = d;
  }
  /* SetCollisionPositions (:211-226): only the first agent's lane is set to 0 */
  for (int i = 0; i < A; ++i)
    for (int j = i + 1; j < A; ++j)
      if (gwo_check_collision(out->x[i], out->last[i], out->lane[i], out->x[j], out->last[j],
                              out->lane[j])) {
        const int c = imax(out->x[i], out->x[j]);
        out->x[i] = c;
        out->x[j] = c;
        out->lane[i] = 0;
        break;
      }
  /* rewards, automata, terminal flags (:58-86) */
  for (int i = 0; i < A; ++i) {
    double* r = rewards + i * V;
    for (int v = 0; v < V; ++v) r[v] = 0.0;
    const uint32_t bits = labels_of(cfg, out, i);
    if (labels) labels[i] = bits;
    if (in->term[i]) {
      out->term[i] = 1; /* :59-62, reward slot defined as 0 here */
      continue;
    }
    gw_slot slots[32];
    const int R = build_slots(cfg, i, slots);
    for (int s = 0; s < R; ++s)
      r[cfg->rules[slots[s].rule].priority] += rule_step(cfg, &slots[s], bits, &out->q[i][s], &out->viol[i][s]);
    /* GetActionCost (:120-132): two separate += on a zero vector, then added */
    const int d = cfg->action_map[actions[i]];
    double cost = 0.0;
    cost += -(double)abs(d - 1) * cfg->speed_deviation_weight;
    const int dd = d - in->last[i];
    cost += -(double)(dd * dd) * cfg->acceleration_weight;
    r[V - 1] += cost;
    out->term[i] = (bits & 1u) || (in->depth + 1 >= cfg->terminal_depth);
  }
  out->depth = in->depth + 1;
}

/* GridWorldState::GetTerminalReward, gridworld/grid_world_state.cpp:141-152 */
static void gw_terminal_reward(const gwo_config* cfg, const gw_state* s, double* rewards) {
  const int A = cfg->n_agents, V = cfg->reward_vec_size;
  for (int i = 0; i < A; ++i) {
    double* r = rewards + i * V;
    for (int v = 0; v < V; ++v) r[v] = 0.0;
    gw_slot slots[32];
    const int R = build_slots(cfg, i, slots);
    for (int k = 0; k < R; ++k) {
      const gwo_rule* rule = &cfg->rules[slots[k].rule];
      r[rule->priority] += rule->accepting[s->q[i][k]] ? 0.0 : (double)rule->weight;
    }
  }
}

static void gw_init_state(const gwo_config* cfg, const int32_t* agents0, const uint8_t* term0,
                          int32_t depth0, const uint8_t* q0, gw_state* s) {
  memset(s, 0, sizeof(*s));
  const int A = cfg->n_agents;
  const int R = gwo_rules_per_agent(cfg);
  for (int i = 0; i < A; ++i) {
    s->x[i] = agents0[4 * i];
    s->last[i] = agents0[4 * i + 1];
    s->lane[i] = agents0[4 * i + 2];
    s->init_lane[i] = agents0[4 * i + 3];
    s->term[i] = term0[i] != 0;
    gw_slot slots[32];
    build_slots(cfg, i, slots);
    for (int k = 0; k < R; ++k)
      s->q[i][k] = q0 ? q0[i * R + k] : (uint8_t)cfg->rules[slots[k].rule].init_state;
  }
  s->depth = depth0;
}

int32_t gwo_replay(GWO_REPLAY_ARGS) {
  const int A = cfg->n_agents, V = cfg->reward_vec_size, R = gwo_rules_per_agent(cfg);
  gw_state s, n;
  gw_init_state(cfg, agents0, term0, depth0, q0, &s);
  int t = 0;
  for (; t < T && !s.term[0]; ++t) {
    gw_execute(cfg, &s, actions + t * A, &n, rewards_out + (size_t)t * A * V, labels_out + t * A);
    s = n;
    for (int a = 0; a < A; ++a) {
      int32_t* o = agents_out + ((size_t)t * A + a) * 4;
      o[0] = s.x[a];
      o[1] = s.last[a];
      o[2] = s.lane[a];
      o[3] = s.init_lane[a];
      term_out[t * A + a] = s.term[a];
      for (int k = 0; k < R; ++k) {
        q_out[((size_t)t * A + a) * R + k] = s.q[a][k];
        viol_out[((size_t)t * A + a) * R + k] = s.viol[a][k];
      }
    }
  }
  gw_terminal_reward(cfg, &s, term_reward_out);
  return t;
}

/* the rollout loop of mvmcts::RandomHeuristic (un-vendored):
 * uniform actions in [0, GetNumActions(i)) until terminal or max_steps, discounted
 * vector return, plus discounted GetTerminalReward(). */
static int rollout_one(const gwo_config* cfg, const gw_state* root, uint64_t seed, uint64_t r,
                       int max_steps, double gamma, int first_exp, double* ret /*[A][V]*/,
                       gw_state* final_state, uint64_t* agent_steps) {
  const int A = cfg->n_agents, V = cfg->reward_vec_size;
  gw_state s = *root, n;
  double rew[GWO_MAX_AGENTS * 8], tr[GWO_MAX_AGENTS * 8];
  for (int i = 0; i < A * V; ++i) ret[i] = 0.0;
  double disc = first_exp ? gamma : 1.0;
  int k = 0;
  while (!s.term[0] && k < max_steps) {
    uint8_t act[GWO_MAX_AGENTS];
    for (int a = 0; a < A; ++a) {
      const uint32_t na = s.term[a] ? 1u : (uint32_t)cfg->n_actions; /* :154-159 */
      act[a] = (uint8_t)oracle_action_draw(seed, r, (uint32_t)k, (uint32_t)a, na);
      if (!s.term[a]) ++*agent_steps;
    }
    gw_execute(cfg, &s, act, &n, rew, NULL);
    for (int i = 0; i < A * V; ++i) ret[i] += disc * rew[i];
    disc *= gamma;
    s = n;
    ++k;
  }
  gw_terminal_reward(cfg, &s, tr);
  for (int i = 0; i < A * V; ++i) ret[i] += disc * tr[i];
  *final_state = s;
  return k;
}

typedef struct {
  const gwo_config* cfg;
  gw_state root;
  uint64_t seed, begin, end;
  int max_steps, first_exp;
  double gamma;
  double ret[GWO_MAX_AGENTS * 8];
  uint64_t viol[GWO_MAX_AGENTS * 32];
  uint64_t steps;
} gw_job;

static void* gw_worker(void* p) {
  gw_job* j = (gw_job*)p;
  const int A = j->cfg->n_agents, V = j->cfg->reward_vec_size, R = gwo_rules_per_agent(j->cfg);
  double ret[GWO_MAX_AGENTS * 8];
  gw_state fin;
  for (uint64_t r = j->begin; r < j->end; ++r) {
    rollout_one(j->cfg, &j->root, j->seed, r, j->max_steps, j->gamma, j->first_exp, ret, &fin,
                &j->steps);
    for (int i = 0; i < A * V; ++i) j->ret[i] += ret[i];
    for (int a = 0; a < A; ++a)
      for (int k = 0; k < R; ++k) j->viol[a * R + k] += fin.viol[a][k];
  }
  return NULL;
}

int32_t gwo_rollout(GWO_ROLLOUT_ARGS) {
  const int A = cfg->n_agents, V = cfg->reward_vec_size, R = gwo_rules_per_agent(cfg);
  const int P = n_threads < 1 ? 1 : (n_threads > 256 ? 256 : n_threads);
  gw_job* jobs = (gw_job*)calloc((size_t)P, sizeof(gw_job));
  pthread_t* th = (pthread_t*)calloc((size_t)P, sizeof(pthread_t));
  for (int p = 0; p < P; ++p) {
    jobs[p].cfg = cfg;
    gw_init_state(cfg, agents0, term0, depth0, q0, &jobs[p].root);
    jobs[p].seed = seed;
    jobs[p].begin = rollout_begin + n_rollouts * (uint64_t)p / (uint64_t)P;
    jobs[p].end = rollout_begin + n_rollouts * (uint64_t)(p + 1) / (uint64_t)P;
    jobs[p].max_steps = max_steps;
    jobs[p].gamma = gamma;
    jobs[p].first_exp = first_discount_exp;
    if (P == 1)
      gw_worker(&jobs[p]);
    else
      pthread_create(&th[p], NULL, gw_worker, &jobs[p]);
  }
  if (P > 1)
    for (int p = 0; p < P; ++p) pthread_join(th[p], NULL);
  for (int i = 0; i < A * V; ++i) {
    returns_sum[i] = 0.0;
    for (int p = 0; p < P; ++p) returns_sum[i] += jobs[p].ret[i];
  }
  for (int i = 0; i < A * R; ++i) {
    viol_sum[i] = 0;
    for (int p = 0; p < P; ++p) viol_sum[i] += jobs[p].viol[i];
  }
  *agent_steps = 0;
  for (int p = 0; p < P; ++p) *agent_steps += jobs[p].steps;
  free(jobs);
  free(th);
  return 0;
}

int32_t gwo_rollout_detail(const gwo_config* cfg, const int32_t* agents0, const uint8_t* term0,
                           int32_t depth0, const uint8_t* q0, uint64_t seed,
                           uint64_t rollout_begin, uint64_t n_rollouts, int32_t max_steps,
                           double gamma, int32_t first_discount_exp, double* returns,
                           uint32_t* viol, uint8_t* q_final, int32_t* agents_final,
                           int32_t* steps) {
  const int A = cfg->n_agents, V = cfg->reward_vec_size, R = gwo_rules_per_agent(cfg);
  gw_state root, fin;
  gw_init_state(cfg, agents0, term0, depth0, q0, &root);
  uint64_t dummy = 0;
  for (uint64_t i = 0; i < n_rollouts; ++i) {
    steps[i] = rollout_one(cfg, &root, seed, rollout_begin + i, max_steps, gamma,
                           first_discount_exp, returns + i * A * V, &fin, &dummy);
    for (int a = 0; a < A; ++a) {
      for (int k = 0; k < R; ++k) {
        viol[(i * A + a) * R + k] = fin.viol[a][k];
        q_final[(i * A + a) * R + k] = fin.q[a][k];
      }
      int32_t* o = agents_final + (i * A + a) * 4;
      o[0] = fin.x[a];
      o[1] = fin.last[a];
      o[2] = fin.lane[a];
      o[3] = fin.init_lane[a];
    }
  }
  return 0;
}

void gwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  oracle_philox4x32_10(ctr, key, out);
}
