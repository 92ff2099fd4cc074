// TEST INFRASTRUCTURE -- CPU restatement of the continuous-world hot path (see the
// header of oracle/cw_oracle.h for scope, citations and the "parity unpinned" note).
// Not part of the product.  All arithmetic in double, the reference's type; compiled with
// -ffp-contract=off so that every operation rounds separately.
#include "oracle/cw_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "oracle/philox.h"

namespace {

struct Slot {
  int rule;
  int other;  // bound other agent, -1: ego-only rule
};

bool rule_is_specific(const cwo_rule& r) {
  for (int a = 0; a < r.n_aps; ++a)
    if (r.ap_label[a] >= 8) return true;
  return false;
}

// Rule states of MCTS agent i: common rules in order, agent-specific ones once per other
// world agent in ascending order (BehaviorRulesMcts::MakeRuleStates with all agents new,
// /root/reference/src/behavior_rules_mcts.hpp:295-322).
int build_slots(const cwo_config& c, int agent, Slot* slots) {
  int n = 0;
  for (int k = 0; k < c.n_rules; ++k) {
    if (rule_is_specific(c.rules[k])) {
      for (int j = 0; j < c.n_agents; ++j)
        if (j != agent) slots[n++] = {k, j};
    } else {
      slots[n++] = {k, -1};
    }
  }
  return n;
}

template <class T>
struct Frenet {
  int lane;  // -1: on no lane
  T s, lat;
};

template <class T>
struct State {
  T x[CWO_MAX_AGENTS], y[CWO_MAX_AGENTS], th[CWO_MAX_AGENTS], v[CWO_MAX_AGENTS];
  uint8_t flags[CWO_MAX_AGENTS];
  int horizon;
  bool terminal;
  uint8_t q[CWO_MAX_AGENTS][CWO_MAX_SLOTS];
  uint32_t viol[CWO_MAX_AGENTS][CWO_MAX_SLOTS];
};

// Pre-processed scenario in the arithmetic type T.
template <class T>
struct World {
  const cwo_config* c;
  int A, M, V, R;
  T dt;
  int n_sub;
  T h_sub, h_last;
  struct Seg {
    T x0, y0, dx, dy, inv_len2, len, s0;
  };
  std::vector<Seg> segs[CWO_MAX_LANES];
  T lane_len[CWO_MAX_LANES];
  struct Edge {
    T ya, yb, xa, m;
  };
  std::vector<Edge> edges;
  T kappa[CWO_MAX_AGENTS][CWO_MAX_PRIMS];  // tan(delta) / wheel_base
  double margin;                            // running min |lhs - rhs| of threshold comparisons

  explicit World(const cwo_config* cfg) : c(cfg) {
    A = c->n_agents;
    M = c->n_mcts_agents;
    V = c->reward_vector_size;
    Slot s[CWO_MAX_SLOTS];
    R = M > 0 ? build_slots(*c, 0, s) : 0;
    // parameters are used as the doubles they are (the reference computes in double
    // throughout, src/rules_mcts_state.cpp:204-264)
    dt = static_cast<T>(c->prediction_time_span);
    // BehaviorMPContinuousActions::Plan (bark, un-vendored): ceil(dt / integration_dt)
    // explicit-Euler sub-steps, the last one takes the remainder.
    n_sub = static_cast<int>(std::ceil(c->prediction_time_span / c->integration_time_delta - 1e-9));
    if (n_sub < 1) n_sub = 1;
    h_sub = static_cast<T>(c->integration_time_delta);
    h_last = static_cast<T>(c->prediction_time_span -
                                               (n_sub - 1) * c->integration_time_delta);
    for (int l = 0; l < c->n_lanes; ++l) {
      const cwo_lane& L = c->lanes[l];
      T s0 = 0;
      for (int k = 0; k + 1 < L.n_pts; ++k) {
        Seg g;
        g.x0 = static_cast<T>(L.pts[k][0]);
        g.y0 = static_cast<T>(L.pts[k][1]);
        const T x1 = static_cast<T>(L.pts[k + 1][0]);
        const T y1 = static_cast<T>(L.pts[k + 1][1]);
        g.dx = x1 - g.x0;
        g.dy = y1 - g.y0;
        const T l2 = g.dx * g.dx + g.dy * g.dy;
        g.inv_len2 = T(1) / l2;
        g.len = std::sqrt(l2);
        g.s0 = s0;
        s0 += g.len;
        segs[l].push_back(g);
      }
      lane_len[l] = s0;
    }
    for (int e = 0; e < c->n_road; ++e) {
      const int f = (e + 1) % c->n_road;
      Edge g;
      g.xa = static_cast<T>(c->road[e][0]);
      g.ya = static_cast<T>(c->road[e][1]);
      const T xb = static_cast<T>(c->road[f][0]);
      g.yb = static_cast<T>(c->road[f][1]);
      g.m = (g.yb != g.ya) ? (xb - g.xa) / (g.yb - g.ya) : T(0);
      edges.push_back(g);
    }
    for (int i = 0; i < A; ++i)
      for (int p = 0; p < c->agents[i].n_prims; ++p)
        kappa[i][p] = static_cast<T>(std::tan(c->agents[i].prim_delta[p]) / c->wheel_base);
    margin = std::numeric_limits<double>::infinity();
  }

  // threshold comparisons record how close they were to flipping
  bool lt(T a, T b) { note(a, b); return a < b; }
  bool le(T a, T b) { note(a, b); return a <= b; }
  bool gt(T a, T b) { note(a, b); return a > b; }
  bool ge(T a, T b) { note(a, b); return a >= b; }
  void note(T a, T b) {
    const double m = std::fabs(static_cast<double>(a) - static_cast<double>(b));
    if (m < margin) margin = m;
  }

  T front(int i) const { return static_cast<T>(c->agents[i].front); }
  T rear(int i) const { return static_cast<T>(c->agents[i].rear); }
  T hw(int i) const { return static_cast<T>(c->agents[i].half_width); }

  // ---- single-track kinematics, explicit Euler sub-steps (bark SingleTrackModel:
  //      x' = v cos(theta), y' = v sin(theta), theta' = v tan(delta) / wheel_base, v' = a)
  void integrate(State<T>& s, int i, int prim) const {
    const T a = static_cast<T>(c->agents[i].prim_acc[prim]);
    const T k = static_cast<T>(kappa[i][prim]);
    T x = s.x[i], y = s.y[i], th = s.th[i], v = s.v[i];
    for (int n = 0; n < n_sub; ++n) {
      const T h = (n == n_sub - 1) ? h_last : h_sub;
      const T cs = std::cos(th), sn = std::sin(th);
      const T nx = x + h * (v * cs);
      const T ny = y + h * (v * sn);
      const T nth = th + h * (v * k);
      const T nv = v + h * a;
      x = nx; y = ny; th = nth; v = nv;
    }
    s.x[i] = x; s.y[i] = y; s.th[i] = th; s.v[i] = v;
  }

  // ---- Frenet projection on the lane centre lines (bark FrenetPosition /
  //      RoadCorridor::GetCurrentLaneCorridor, call site src/rules_mcts_state.cpp:232-238)
  Frenet<T> frenet(const State<T>& s, int i) {
    Frenet<T> out;
    out.lane = -1;
    out.s = 0;
    out.lat = 0;
    for (int l = 0; l < c->n_lanes; ++l) {
      T best = std::numeric_limits<T>::infinity(), bt = 0, btu = 0, bcross = 0;
      int bk = 0;
      const int ns = static_cast<int>(segs[l].size());
      for (int k = 0; k < ns; ++k) {
        const Seg& g = segs[l][k];
        const T wx = s.x[i] - g.x0, wy = s.y[i] - g.y0;
        const T tu = (wx * g.dx + wy * g.dy) * g.inv_len2;
        const T t = std::min(std::max(tu, T(0)), T(1));
        const T ex = wx - t * g.dx, ey = wy - t * g.dy;
        const T d2 = ex * ex + ey * ey;
        if (d2 < best) {
          best = d2; bk = k; bt = t; btu = tu;
          bcross = g.dx * wy - g.dy * wx;
        }
      }
      const T hwl = static_cast<T>(c->lanes[l].half_width);
      const bool interior = !(bk == 0 && lt(btu, T(0))) && !(bk == ns - 1 && gt(btu, T(1)));
      if (interior && le(best, hwl * hwl)) {
        out.lane = l;
        out.s = segs[l][bk].s0 + bt * segs[l][bk].len;
        const T d = std::sqrt(best);
        out.lat = bcross < 0 ? -d : d;
        return out;
      }
    }
    return out;
  }

  // ---- rectangle-rectangle intersection (separating axes); EvaluatorCollisionEgoAgent
  //      tests the agent's polygon against every other agent's polygon.
  bool rect_overlap(const State<T>& s, int i, int j) {
    const T ci = std::cos(s.th[i]), si = std::sin(s.th[i]);
    const T cj = std::cos(s.th[j]), sj = std::sin(s.th[j]);
    const T oi = (front(i) - rear(i)) * T(0.5), oj = (front(j) - rear(j)) * T(0.5);
    const T eix = (front(i) + rear(i)) * T(0.5), ejx = (front(j) + rear(j)) * T(0.5);
    const T eiy = hw(i), ejy = hw(j);
    const T dx = (s.x[j] + oj * cj) - (s.x[i] + oi * ci);
    const T dy = (s.y[j] + oj * sj) - (s.y[i] + oi * si);
    // |u_i.u_j| = |w_i.w_j| = |cos(dth)|, |u_i.w_j| = |w_i.u_j| = |sin(dth)|
    const T cd = std::fabs(ci * cj + si * sj), sd = std::fabs(si * cj - ci * sj);
    bool hit = true;
    hit = le(std::fabs(dx * ci + dy * si), eix + ejx * cd + ejy * sd) && hit;   // axis u_i
    hit = le(std::fabs(-dx * si + dy * ci), eiy + ejx * sd + ejy * cd) && hit;  // axis w_i
    hit = le(std::fabs(dx * cj + dy * sj), ejx + eix * cd + eiy * sd) && hit;   // axis u_j
    hit = le(std::fabs(-dx * sj + dy * cj), ejy + eix * sd + eiy * cd) && hit;  // axis w_j
    return hit;
  }

  bool collision_ego(const State<T>& s, int i) {
    bool hit = false;
    for (int j = 0; j < A; ++j) {
      if (j == i || !(s.flags[j] & CWO_PRESENT)) continue;
      hit = rect_overlap(s, i, j) || hit;
    }
    return hit;
  }

  void corners(const State<T>& s, int i, T (&cx)[4], T (&cy)[4]) const {
    const T cs = std::cos(s.th[i]), sn = std::sin(s.th[i]);
    const T lx[4] = {front(i), front(i), -rear(i), -rear(i)};
    const T ly[4] = {hw(i), -hw(i), -hw(i), hw(i)};
    for (int k = 0; k < 4; ++k) {
      cx[k] = s.x[i] + (lx[k] * cs - ly[k] * sn);
      cy[k] = s.y[i] + (lx[k] * sn + ly[k] * cs);
    }
  }

  bool point_in_road(T px, T py) {
    bool inside = false;
    for (const Edge& g : edges) {
      if (gt(g.ya, py) != gt(g.yb, py)) {
        if (lt(px, g.m * (py - g.ya) + g.xa)) inside = !inside;
      }
    }
    return inside;
  }

  // EvaluatorDrivableArea: the agent polygon is not within the road polygon.  Restated as
  // "a corner lies outside, or a road vertex lies strictly inside the rectangle".
  bool out_of_map(const State<T>& s, int i) {
    T cx[4], cy[4];
    corners(s, i, cx, cy);
    bool out = false;
    for (int k = 0; k < 4; ++k) out = !point_in_road(cx[k], cy[k]) || out;
    const T cs = std::cos(s.th[i]), sn = std::sin(s.th[i]);
    for (const Edge& g : edges) {
      const T dx = g.xa - s.x[i], dy = g.ya - s.y[i];
      const T lx = dx * cs + dy * sn, ly = -dx * sn + dy * cs;
      const bool in = gt(lx, -rear(i)) && lt(lx, front(i)) && lt(std::fabs(ly), hw(i));
      out = in || out;
    }
    return out;
  }

  // GoalDefinitionPolygon / GoalDefinitionStateLimits: agent shape within the goal
  // rectangle (and heading within limits).
  bool at_goal(const State<T>& s, int i) {
    const cwo_agent& ag = c->agents[i];
    if (!ag.has_goal) return false;
    T cx[4], cy[4];
    corners(s, i, cx, cy);
    bool in = true;
    const T g0 = static_cast<T>(ag.goal[0]), g1 = static_cast<T>(ag.goal[1]);
    const T g2 = static_cast<T>(ag.goal[2]), g3 = static_cast<T>(ag.goal[3]);
    for (int k = 0; k < 4; ++k)
      in = ge(cx[k], g0) && le(cx[k], g1) && ge(cy[k], g2) && le(cy[k], g3) && in;
    in = ge(s.th[i], static_cast<T>(ag.goal[4])) &&
         le(s.th[i], static_cast<T>(ag.goal[5])) && in;
    return in;
  }

  // gap from i to j along i's corridor (own lane, then its successor); false if j is not there
  bool gap(const Frenet<T>* f, int i, int j, T* g) const {
    if (f[i].lane < 0 || f[j].lane < 0) return false;
    if (f[j].lane == f[i].lane) {
      *g = f[j].s - f[i].s;
      return true;
    }
    if (c->lanes[f[i].lane].successor == f[j].lane) {
      *g = (lane_len[f[i].lane] - f[i].s) + f[j].s;
      return true;
    }
    return false;
  }

  // labels of MCTS agent i: word 0 ego bits by kind, words 1..3 masks over j (kinds 8,9,10)
  void labels(const State<T>& s, const Frenet<T>* f, int i, bool coll, bool oom, bool goal,
              uint32_t (&w)[4]) {
    w[0] = w[1] = w[2] = w[3] = 0;
    if (coll) w[0] |= 1u << CWO_LABEL_COLLISION_EGO;
    if (oom) w[0] |= 1u << CWO_LABEL_COLLISION_CORRIDOR;
    if (goal) w[0] |= 1u << CWO_LABEL_GOAL_REACHED;
    // preceding agent = smallest positive gap in i's corridor
    int jf = -1;
    T gf = 0;
    for (int j = 0; j < A; ++j) {
      if (j == i || !(s.flags[j] & CWO_PRESENT)) continue;
      T g;
      if (!gap(f, i, j, &g)) continue;
      if (!gt(g, T(0))) continue;
      if (jf < 0 || lt(g, gf)) {
        jf = j;
        gf = g;
      }
    }
    if (jf >= 0) w[2] |= 1u << jf;
    // safe distance to the preceding agent (SafeDistanceLabelFunction, restated as
    // d_safe = v_r*delta + v_r^2/(2|a_r|) - v_f^2/(2|a_f|), bumper-to-bumper gap)
    bool safe = true;
    if (jf >= 0) {
      const T dist = gf - (front(i) + rear(jf));
      const T vr = s.v[i], vf = s.v[jf];
      const T delta = static_cast<T>(c->sd_reaction_time);
      const T ar = static_cast<T>(1.0 / (2.0 * std::fabs(c->sd_max_decel_rear)));
      const T af = static_cast<T>(1.0 / (2.0 * std::fabs(c->sd_max_decel_front)));
      const T dsafe = vr * delta + (vr * vr) * ar - (vf * vf) * af;
      safe = ge(dist, dsafe);
    }
    if (safe) w[0] |= 1u << CWO_LABEL_SAFE_DISTANCE_FRONT;
    auto beyond = [&](int k) {
      if (f[k].lane < 0) return false;
      return ge(f[k].s, static_cast<T>(c->lanes[f[k].lane].s_point));
    };
    if (beyond(i)) w[0] |= 1u << CWO_LABEL_MERGED_E;
    if (f[i].lane >= 0 && (c->lanes[f[i].lane].flags & 1)) w[0] |= 1u << CWO_LABEL_RIGHTMOST_LANE;
    if (f[i].lane >= 0) w[0] |= 1u << CWO_LABEL_ON_ROAD;
    const T nd = static_cast<T>(c->near_distance);
    for (int j = 0; j < A; ++j) {
      if (j == i || !(s.flags[j] & CWO_PRESENT)) continue;
      if (beyond(j)) w[1] |= 1u << j;
      const T dx = s.x[j] - s.x[i], dy = s.y[j] - s.y[i];
      if (lt(dx * dx + dy * dy, nd * nd)) w[3] |= 1u << j;
    }
  }

  static int label_value(const uint32_t (&w)[4], int kind, int other) {
    if (kind < 8) return (w[0] >> kind) & 1u;
    return (w[kind - 7] >> other) & 1u;
  }

  // one automaton step (contract of ltl::RuleMonitor::Evaluate, src/rules_mcts_state.cpp:179-182)
  T rule_step(const Slot& slot, const uint32_t (&w)[4], uint8_t* q, uint32_t* viol) const {
    const cwo_rule& r = c->rules[slot.rule];
    int letter = 0;
    for (int k = 0; k < r.n_aps; ++k)
      letter |= label_value(w, r.ap_label[k], slot.other) << (r.n_aps - 1 - k);
    const uint8_t next = r.trans[(*q) * (1 << r.n_aps) + letter];
    if (next == CWO_VIOLATION) {
      ++*viol;
      if (r.reset_on_violation) *q = static_cast<uint8_t>(r.init_state);
      return static_cast<T>(r.weight);
    }
    *q = next;
    return T(0);
  }

  // CheckTerminal (src/rules_mcts_state.cpp:266-285)
  bool check_terminal(const State<T>& s) {
    if (!(s.flags[0] & CWO_PRESENT)) return true;
    const bool horizon_reached = (s.horizon == 0);
    const bool oom = out_of_map(s, 0);
    const bool coll = collision_ego(s, 0);
    const bool goal = at_goal(s, 0);
    return horizon_reached || oom || coll || goal;
  }

  // MvmctsState::Execute (src/rules_mcts_state.cpp:57-128)
  void execute(const State<T>& in, const uint8_t* actions, State<T>& out, T* rewards /*[M][V]*/,
               uint32_t* labels_out /*[M][4] or null*/, Frenet<T>* frenet_out /*[A] or null*/,
               uint64_t* agent_steps) {
    out = in;
    // Predict (:69-71): valid agents follow their primitive, invalid (collided) ones are frozen
    for (int i = 0; i < A; ++i) {
      if (!(in.flags[i] & CWO_PRESENT) || !(in.flags[i] & CWO_VALID)) continue;
      const int prim = (i < M) ? actions[i] : 0;
      integrate(out, i, prim);
      if (agent_steps) ++*agent_steps;
    }
    out.horizon = in.horizon - 1;
    // World::Step -> RemoveInvalidAgents (bark, un-vendored; parameter
    // "World::remove_agents_out_of_map", default false): agents whose shape is no longer
    // within the road polygon leave the world
    if (c->remove_agents_out_of_map)
      for (int i = 0; i < A; ++i)
        if ((out.flags[i] & CWO_PRESENT) && out_of_map(out, i)) out.flags[i] = 0;
    Frenet<T> f[CWO_MAX_AGENTS];
    for (int i = 0; i < A; ++i) {
      f[i].lane = -1; f[i].s = 0; f[i].lat = 0;
      if (out.flags[i] & CWO_PRESENT) f[i] = frenet(out, i);
      if (frenet_out) frenet_out[i] = f[i];
    }
    const T gamma = static_cast<T>(c->discount_factor);
    const T vdes = static_cast<T>(c->desired_velocity);
    bool ego_coll = false, ego_oom = false, ego_goal = false;
    for (int ai = 0; ai < M; ++ai) {
      T* r = rewards + ai * V;
      for (int v = 0; v < V; ++v) r[v] = 0;
      uint32_t w[4] = {0, 0, 0, 0};
      if ((out.flags[ai] & CWO_PRESENT) && (in.flags[ai] & CWO_VALID)) {
        const bool coll = collision_ego(out, ai);
        const bool oom = out_of_map(out, ai);
        const bool goal = at_goal(out, ai);
        if (ai == 0) { ego_coll = coll; ego_oom = oom; ego_goal = goal; }
        r[0] += (coll ? T(1) : T(0)) * static_cast<T>(c->collision_weight);   // :91-92
        r[0] += (oom ? T(1) : T(0)) * static_cast<T>(c->out_of_map_weight);   // :94-96
        // GetActionCost (:202-248)
        {
          const T a = (out.v[ai] - in.v[ai]) / dt;
          T cost = static_cast<T>(c->acceleration_weight) * a * a * dt;
          const T theta_dot = (out.th[ai] - in.th[ai]) / dt;
          const T avg_vel = T(0.5) * a * dt + in.v[ai];
          const T a_lat = theta_dot * avg_vel;
          cost += static_cast<T>(c->radial_acceleration_weight) * a_lat * a_lat * dt;
          cost += static_cast<T>(c->desired_velocity_weight) *
                  std::fabs(in.v[ai] - vdes) * dt;
          if (f[ai].lane >= 0)
            cost += static_cast<T>(c->lane_center_weight) * std::fabs(f[ai].lat) * dt;
          r[V - 1] += cost;
        }
        // PotentialReward (:250-264)
        {
          const T pw = static_cast<T>(c->potential_weight);
          const T pot_new = -pw * dt * std::fabs(out.v[ai] - vdes);
          const T pot_cur = -pw * dt * std::fabs(in.v[ai] - vdes);
          r[V - 1] += gamma * pot_new - pot_cur;
        }
        labels(out, f, ai, coll, oom, goal, w);
        if (!(c->use_rule_reward_for_ego_only && ai != 0)) {  // :104-108
          T rr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          Slot slots[CWO_MAX_SLOTS];
          const int nr = build_slots(*c, ai, slots);
          for (int k = 0; k < nr; ++k)
            rr[c->rules[slots[k].rule].priority] += rule_step(slots[k], w, &out.q[ai][k], &out.viol[ai][k]);
          for (int v = 0; v < V; ++v) r[v] += rr[v];
        }
        if (goal) r[V - 1] += static_cast<T>(c->goal_reward);  // :110-113
        if (coll) out.flags[ai] &= static_cast<uint8_t>(~CWO_VALID);               // :116-120
      } else if ((in.flags[ai] & CWO_PRESENT) && !(out.flags[ai] & CWO_PRESENT)) {
        r[0] += static_cast<T>(c->out_of_map_weight);  // left the map during this step, :121-125
      }
      if (labels_out) std::memcpy(labels_out + ai * 4, w, sizeof(w));
    }
    // ego terminal test of the new state (ctor -> CheckTerminal, :52, :266-285); when the
    // ego was evaluated above the three predicates are the same computations
    if (!(out.flags[0] & CWO_PRESENT)) {
      out.terminal = true;
    } else if (in.flags[0] & CWO_VALID) {
      out.terminal = (out.horizon == 0) || ego_oom || ego_coll || ego_goal;
    } else {
      out.terminal = check_terminal(out);
    }
  }

  // MvmctsState::EvaluateRules() (:186-197): labels of the state as it is, one transition per
  // rule state of every MCTS agent that is present and VALID; no other reward term
  void evaluate_rules(State<T>& s, T* rewards, uint32_t* labels_out) {
    Frenet<T> f[CWO_MAX_AGENTS];
    for (int i = 0; i < A; ++i) {
      f[i].lane = -1; f[i].s = 0; f[i].lat = 0;
      if (s.flags[i] & CWO_PRESENT) f[i] = frenet(s, i);
    }
    for (int ai = 0; ai < M; ++ai) {
      T* r = rewards + ai * V;
      for (int v = 0; v < V; ++v) r[v] = 0;
      uint32_t w[4] = {0, 0, 0, 0};
      if ((s.flags[ai] & CWO_PRESENT) && (s.flags[ai] & CWO_VALID)) {
        const bool coll = collision_ego(s, ai);
        const bool oom = out_of_map(s, ai);
        const bool goal = at_goal(s, ai);
        labels(s, f, ai, coll, oom, goal, w);
        Slot slots[CWO_MAX_SLOTS];
        const int nr = build_slots(*c, ai, slots);
        for (int k = 0; k < nr; ++k)
          r[c->rules[slots[k].rule].priority] += rule_step(slots[k], w, &s.q[ai][k], &s.viol[ai][k]);
      }
      if (labels_out) std::memcpy(labels_out + ai * 4, w, sizeof(w));
    }
  }

  // GetTerminalReward (:148-164)
  void terminal_reward(const State<T>& s, T* rewards) const {
    for (int ai = 0; ai < M; ++ai) {
      T* r = rewards + ai * V;
      for (int v = 0; v < V; ++v) r[v] = 0;
      if (!(s.flags[ai] & CWO_PRESENT)) continue;
      Slot slots[CWO_MAX_SLOTS];
      const int nr = build_slots(*c, ai, slots);
      for (int k = 0; k < nr; ++k) {
        const cwo_rule& rule = c->rules[slots[k].rule];
        r[rule.priority] += rule.accepting[s.q[ai][k]] ? T(0) : static_cast<T>(rule.weight);
      }
    }
  }

  void init_state(const double* states0, const uint8_t* flags0, int horizon0, const uint8_t* q0,
                  State<T>& s) {
    std::memset(&s, 0, sizeof(s));
    for (int i = 0; i < A; ++i) {
      s.x[i] = static_cast<T>(states0[4 * i]);
      s.y[i] = static_cast<T>(states0[4 * i + 1]);
      s.th[i] = static_cast<T>(states0[4 * i + 2]);
      s.v[i] = static_cast<T>(states0[4 * i + 3]);
      s.flags[i] = flags0[i];
    }
    s.horizon = horizon0;
    for (int ai = 0; ai < M; ++ai) {
      Slot slots[CWO_MAX_SLOTS];
      const int nr = build_slots(*c, ai, slots);
      for (int k = 0; k < nr; ++k)
        s.q[ai][k] = q0 ? q0[ai * R + k] : static_cast<uint8_t>(c->rules[slots[k].rule].init_state);
    }
    s.terminal = check_terminal(s);
  }

  // GetNumActions (:129-141)
  uint32_t num_actions(const State<T>& s, int i) const {
    return ((s.flags[i] & CWO_PRESENT) && (s.flags[i] & CWO_VALID)) ? c->agents[i].n_prims : 1;
  }

  // mvmcts::RandomHeuristic loop (un-vendored)
  int rollout_one(const State<T>& root, uint64_t seed, uint64_t rid, int max_steps, T gamma,
                  int first_exp, T* ret /*[M][V]*/, State<T>& fin, uint64_t* agent_steps) {
    State<T> s = root, n;
    T rew[CWO_MAX_AGENTS * 8], tr[CWO_MAX_AGENTS * 8];
    for (int i = 0; i < M * V; ++i) ret[i] = 0;
    T disc = first_exp ? gamma : T(1);
    int k = 0;
    while (!s.terminal && k < max_steps) {
      uint8_t act[CWO_MAX_AGENTS];
      for (int a = 0; a < M; ++a)
        act[a] = static_cast<uint8_t>(oracle_action_draw(seed, rid, k, a, num_actions(s, a)));
      execute(s, act, n, rew, nullptr, nullptr, agent_steps);
      for (int i = 0; i < M * V; ++i) ret[i] += disc * rew[i];
      disc *= gamma;
      s = n;
      ++k;
    }
    terminal_reward(s, tr);
    for (int i = 0; i < M * V; ++i) ret[i] += disc * tr[i];
    fin = s;
    return k;
  }
};

template <class T>
int replay_t(const cwo_config* cfg, const double* states0, const uint8_t* flags0, int32_t horizon0,
             const uint8_t* q0, const uint8_t* actions, int32_t Tn, double* states_out,
             uint8_t* flags_out, uint8_t* terminal_out, uint8_t* q_out, uint32_t* viol_out,
             double* rewards_out, uint32_t* labels_out, double* frenet_out, double* margin_out,
             double* term_reward_out) {
  World<T> w(cfg);
  const int A = w.A, M = w.M, V = w.V, R = w.R;
  State<T> s, n;
  w.init_state(states0, flags0, horizon0, q0, s);
  int t = 0;
  for (; t < Tn && !s.terminal; ++t) {
    T rew[CWO_MAX_AGENTS * 8];
    Frenet<T> f[CWO_MAX_AGENTS];
    w.margin = std::numeric_limits<double>::infinity();
    w.execute(s, actions + t * M, n, rew, labels_out + static_cast<size_t>(t) * M * 4, f, nullptr);
    margin_out[t] = w.margin;
    s = n;
    for (int i = 0; i < A; ++i) {
      double* o = states_out + (static_cast<size_t>(t) * A + i) * 4;
      o[0] = s.x[i]; o[1] = s.y[i]; o[2] = s.th[i]; o[3] = s.v[i];
      flags_out[t * A + i] = s.flags[i];
      double* fo = frenet_out + (static_cast<size_t>(t) * A + i) * 3;
      fo[0] = f[i].lane; fo[1] = f[i].s; fo[2] = f[i].lat;
    }
    terminal_out[t] = s.terminal;
    for (int a = 0; a < M; ++a) {
      for (int k = 0; k < R; ++k) {
        q_out[(static_cast<size_t>(t) * M + a) * R + k] = s.q[a][k];
        viol_out[(static_cast<size_t>(t) * M + a) * R + k] = s.viol[a][k];
      }
      for (int v = 0; v < V; ++v)
        rewards_out[(static_cast<size_t>(t) * M + a) * V + v] = rew[a * V + v];
    }
  }
  T tr[CWO_MAX_AGENTS * 8];
  w.terminal_reward(s, tr);
  for (int i = 0; i < M * V; ++i) term_reward_out[i] = tr[i];
  return t;
}

template <class T>
void rollout_range(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                   int32_t horizon0, const uint8_t* q0, uint64_t seed, uint64_t begin, uint64_t end,
                   int max_steps, double gamma, int first_exp, double* ret_sum, uint64_t* viol_sum,
                   uint64_t* agent_steps, double* returns, uint32_t* viol, uint8_t* q_final,
                   double* states_final, uint8_t* flags_final, int32_t* steps, double* margin) {
  World<T> w(cfg);
  const int A = w.A, M = w.M, V = w.V, R = w.R;
  State<T> root, fin;
  w.init_state(states0, flags0, horizon0, q0, root);
  T ret[CWO_MAX_AGENTS * 8];
  for (uint64_t r = begin; r < end; ++r) {
    const uint64_t i = r - begin;
    w.margin = std::numeric_limits<double>::infinity();
    const int k = w.rollout_one(root, seed, r, max_steps, static_cast<T>(gamma), first_exp, ret, fin,
                                agent_steps);
    if (ret_sum)
      for (int j = 0; j < M * V; ++j) ret_sum[j] += ret[j];
    if (viol_sum)
      for (int a = 0; a < M; ++a)
        for (int s = 0; s < R; ++s) viol_sum[a * R + s] += fin.viol[a][s];
    if (returns) {
      for (int j = 0; j < M * V; ++j) returns[i * M * V + j] = ret[j];
      for (int a = 0; a < M; ++a)
        for (int s = 0; s < R; ++s) {
          viol[(i * M + a) * R + s] = fin.viol[a][s];
          q_final[(i * M + a) * R + s] = fin.q[a][s];
        }
      for (int a = 0; a < A; ++a) {
        double* o = states_final + (i * A + a) * 4;
        o[0] = fin.x[a]; o[1] = fin.y[a]; o[2] = fin.th[a]; o[3] = fin.v[a];
        flags_final[i * A + a] = fin.flags[a];
      }
      steps[i] = k;
      margin[i] = w.margin;
    }
  }
}

}  // namespace

extern "C" {

int32_t cwo_rules_per_agent(const cwo_config* cfg) {
  Slot s[CWO_MAX_SLOTS];
  return cfg->n_mcts_agents > 0 ? build_slots(*cfg, 0, s) : 0;
}

int32_t cwo_check_terminal(const cwo_config* cfg, const double* states, const uint8_t* flags,
                           int32_t horizon) {
  World<double> w(cfg);
  State<double> s;
  w.init_state(states, flags, horizon, nullptr, s);
  return s.terminal;
}

int32_t cwo_replay(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                   int32_t horizon0, const uint8_t* q0, const uint8_t* actions, int32_t T,
                   double* states_out, uint8_t* flags_out, uint8_t* terminal_out, uint8_t* q_out,
                   uint32_t* viol_out, double* rewards_out, uint32_t* labels_out,
                   double* frenet_out, double* margin_out, double* term_reward_out) {
  return replay_t<double>(cfg, states0, flags0, horizon0, q0, actions, T, states_out, flags_out,
                          terminal_out, q_out, viol_out, rewards_out, labels_out, frenet_out,
                          margin_out, term_reward_out);
}

int32_t cwo_evaluate_rules(const cwo_config* cfg, const double* states, const uint8_t* flags,
                           int32_t horizon, const uint8_t* q_in, uint8_t* q_out, uint32_t* viol_out,
                           double* rewards_out, uint32_t* labels_out) {
  World<double> w(cfg);
  State<double> s;
  w.init_state(states, flags, horizon, q_in, s);
  double rew[CWO_MAX_AGENTS * 8];
  w.evaluate_rules(s, rew, labels_out);
  for (int a = 0; a < w.M; ++a) {
    for (int k = 0; k < w.R; ++k) {
      q_out[a * w.R + k] = s.q[a][k];
      viol_out[a * w.R + k] = s.viol[a][k];
    }
    for (int v = 0; v < w.V; ++v) rewards_out[a * w.V + v] = rew[a * w.V + v];
  }
  return 0;
}

void cwo_coop_mix(double* returns_sum, int32_t M, int32_t V, double coop_factor) {
  // mvmcts COOP_FACTOR (src/util.cpp:28-29, 0 in every test of the reference; the mixing rule
  // lives in the un-vendored mvmcts and is restated from the survey's recollection):
  // value_i = own_i + coop * sum_{j != i} own_j, per reward component
  if (coop_factor == 0.0) return;
  double own[CWO_MAX_AGENTS * 8];
  for (int i = 0; i < M * V; ++i) own[i] = returns_sum[i];
  for (int a = 0; a < M; ++a)
    for (int v = 0; v < V; ++v) {
      double others = 0.0;
      for (int j = 0; j < M; ++j)
        if (j != a) others += own[j * V + v];
      returns_sum[a * V + v] = own[a * V + v] + coop_factor * others;
    }
}

int32_t cwo_rollout(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                    int32_t horizon0, const uint8_t* q0, uint64_t seed, uint64_t rollout_begin,
                    uint64_t n_rollouts, int32_t max_steps, double gamma,
                    int32_t first_discount_exp, double coop_factor, int32_t n_threads,
                    double* returns_sum, uint64_t* viol_sum, uint64_t* agent_steps) {
  const int M = cfg->n_mcts_agents, V = cfg->reward_vector_size, R = cwo_rules_per_agent(cfg);
  const int P = n_threads < 1 ? 1 : n_threads;
  std::vector<std::vector<double>> rs(P, std::vector<double>(M * V, 0.0));
  std::vector<std::vector<uint64_t>> vs(P, std::vector<uint64_t>(M * R + 1, 0));
  std::vector<uint64_t> st(P, 0);
  std::vector<std::thread> th;
  for (int p = 0; p < P; ++p) {
    const uint64_t b = rollout_begin + n_rollouts * p / P, e = rollout_begin + n_rollouts * (p + 1) / P;
    auto job = [=, &rs, &vs, &st]() {
      // thread-private accumulators (no false sharing between workers), published at the end
      std::vector<double> lr(M * V, 0.0);
      std::vector<uint64_t> lv(M * R + 1, 0);
      uint64_t ls = 0;
      rollout_range<double>(cfg, states0, flags0, horizon0, q0, seed, b, e, max_steps, gamma,
                            first_discount_exp, lr.data(), lv.data(), &ls, nullptr, nullptr,
                            nullptr, nullptr, nullptr, nullptr, nullptr);
      rs[p] = lr;
      vs[p] = lv;
      st[p] = ls;
    };
    if (P == 1) job(); else th.emplace_back(job);
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < M * V; ++i) {
    returns_sum[i] = 0;
    for (int p = 0; p < P; ++p) returns_sum[i] += rs[p][i];
  }
  cwo_coop_mix(returns_sum, M, V, coop_factor);
  for (int i = 0; i < M * R; ++i) {
    viol_sum[i] = 0;
    for (int p = 0; p < P; ++p) viol_sum[i] += vs[p][i];
  }
  *agent_steps = 0;
  for (int p = 0; p < P; ++p) *agent_steps += st[p];
  return 0;
}

int32_t cwo_rollout_detail(const cwo_config* cfg, const double* states0, const uint8_t* flags0,
                           int32_t horizon0, const uint8_t* q0, uint64_t seed,
                           uint64_t rollout_begin, uint64_t n_rollouts, int32_t max_steps,
                           double gamma, int32_t first_discount_exp, double* returns,
                           uint32_t* viol, uint8_t* q_final, double* states_final,
                           uint8_t* flags_final, int32_t* steps, double* margin) {
  uint64_t dummy = 0;
  rollout_range<double>(cfg, states0, flags0, horizon0, q0, seed, rollout_begin,
                        rollout_begin + n_rollouts, max_steps, gamma, first_discount_exp, nullptr,
                        nullptr, &dummy, returns, viol, q_final, states_final, flags_final, steps, margin);
  return 0;
}

}  // extern "C"
