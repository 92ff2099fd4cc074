// Device-side GridWorld step: the arithmetic of GridWorldState::Execute
// (/root/reference/gridworld/grid_world_state.cpp:45-132,211-226) and of the four label
// evaluators (/root/reference/gridworld/label_evaluator/*.cpp), written for one
// thread = one state with every agent in registers.  Semantics: SURVEY.md Appendix A.
#pragma once
#include "common.cuh"

namespace brm {

template <int A>
struct GwRegs {
  int x[A], last[A], lane[A], init_lane[A];
  uint32_t term;   // bit a: agent a terminal
  int depth;
  uint64_t q[A];   // 16 rule slots x 4-bit automaton state
};

// evaluator_label_at_position.cpp:16-28
__device__ __forceinline__ bool gw_at_position(int x, int last, int p) {
  const int prev = x - last;
  return (prev < p && x >= p) || (prev > p && x <= p) || x == p;
}

// check_collision, evaluator_label_collision.cpp:14-31
__device__ __forceinline__ bool gw_check_collision(int ax, int alast, int alane, int bx, int blast,
                                                   int blane) {
  const int o_prev = bx - blast;
  const int upper_o = max(o_prev, bx), lower_o = min(o_prev, bx);
  const int e_prev = ax - alast;
  const int upper_e = max(e_prev, ax), lower_e = min(e_prev, ax);
  const int overlap = min(upper_e, upper_o) - max(lower_e, lower_o);
  const bool moved_onto = (blast != 0 && bx == e_prev) || (alast != 0 && ax == o_prev);
  const bool is_overlapping = (!moved_onto || overlap != 0) && overlap >= 0;
  return (blane == alane) && (is_overlapping || ax == bx);
}

struct GwLabels {
  uint32_t ego;     // bit0 collision, bit1 merged_e
  uint32_t merged;  // bit j: merged_x#j (own bit cleared)
  uint32_t front;   // bit j: in_direct_front_x#j
};

// One Execute.  `sink(i, r)` receives the reward vector of every agent (zeros for
// agents that were already terminal); `on_violation(i, slot)` is called per violated
// rule slot; `on_labels(i, lab)` receives the label valuation seen by agent i.
template <int A, class RewardSink, class ViolSink, class LabelSink>
__device__ __forceinline__ void gw_step(const GwBlob& B, GwRegs<A>& s, const uint32_t (&act)[A],
                                        RewardSink&& sink, ViolSink&& on_violation,
                                        LabelSink&& on_labels) {
  const int mp = B.merging_point;
  const uint32_t term_in = s.term;
  int last_in[A];
  int delta[A];
  // Step (:91-119)
#pragma unroll
  for (int i = 0; i < A; ++i) {
    last_in[i] = s.last[i];
    delta[i] = B.action_map[act[i]];
    if (!((term_in >> i) & 1u)) {
      const int nx = s.x[i] + delta[i];
      s.x[i] = nx;
      s.lane[i] = (nx >= mp) ? 0 : s.init_lane[i];
      s.last[i] = delta[i];
    }
  }
  // SetCollisionPositions (:211-226): sequential, uses already snapped values,
  // only the first agent of a pair is moved to lane 0, first hit per i wins.
#pragma unroll
  for (int i = 0; i < A; ++i) {
    bool done = false;
#pragma unroll
    for (int j = i + 1; j < A; ++j) {
      if (!done && gw_check_collision(s.x[i], s.last[i], s.lane[i], s.x[j], s.last[j], s.lane[j])) {
        const int c = max(s.x[i], s.x[j]);
        s.x[i] = c;
        s.x[j] = c;
        s.lane[i] = 0;
        done = true;
      }
    }
  }
  // shared per-step predicates
  uint32_t atx = 0, merged = 0;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    if (gw_at_position(s.x[i], s.last[i], mp)) atx |= 1u << i;
    if (s.x[i] >= mp && s.x[i] <= B.x_len) merged |= 1u << i;
  }
  uint32_t term_out = 0;
  const bool depth_hit = (s.depth + 1 >= B.terminal_depth);
#pragma unroll
  for (int i = 0; i < A; ++i) {
    GwLabels lab;
    // collision (evaluator_label_collision.cpp:36-54)
    bool coll = false;
#pragma unroll
    for (int j = 0; j < A; ++j) {
      if (j == i) continue;
      coll = coll || (((atx >> i) & 1u) && ((atx >> j) & 1u)) ||
             gw_check_collision(s.x[i], s.last[i], s.lane[i], s.x[j], s.last[j], s.lane[j]);
    }
    // in_direct_front_x#j (evaluator_label_in_direct_front.cpp:15-34)
    uint32_t front = 0;
#pragma unroll
    for (int j = 0; j < A; ++j) {
      if (j == i) continue;
      bool f = (s.lane[j] == s.lane[i]) && (s.x[j] > s.x[i]);
#pragma unroll
      for (int k = 0; k < A; ++k) {
        if (k == i || k == j) continue;
        if (s.x[k] >= s.x[i] && s.x[k] <= s.x[j] && s.lane[k] == s.lane[i]) f = false;
      }
      if (f) front |= 1u << j;
    }
    lab.ego = (coll ? 1u : 0u) | (((merged >> i) & 1u) << 1);
    lab.merged = merged & ~(1u << i);
    lab.front = front;
    on_labels(i, lab);

    double r[BRM_MAX_REWARD];
#pragma unroll
    for (int v = 0; v < BRM_MAX_REWARD; ++v) r[v] = 0.0;
    if ((term_in >> i) & 1u) {
      term_out |= 1u << i;  // :59-62 (reward slot defined as 0)
      sink(i, r);
      continue;
    }
    // automata transit (:76-79)
    // label word of this agent: the slot records say which bit each AP reads (common.cuh)
    const uint32_t word = lab.ego | (lab.merged << 2) | (lab.front << 10);
    for (int sl = 0; sl < B.R; ++sl) {
      const uint32_t bits = B.slot_bits[i][sl], meta = B.slot_meta[i][sl];
      const uint32_t n_aps = (meta >> 8) & 0xffu;
      uint32_t letter = 0;
#pragma unroll
      for (uint32_t k = 0; k < BRM_MAX_APS; ++k) letter |= ((word >> ((bits >> (8 * k)) & 31u)) & 1u) << k;
      const uint32_t cur = static_cast<uint32_t>(s.q[i] >> (4 * sl)) & 15u;
      uint32_t nxt = B.rules[meta & 0xffu].trans[(cur << n_aps) + letter];
      if (nxt == BRM_VIOLATION) {
        on_violation(i, sl);
        nxt = (meta >> 31) ? cur : ((meta >> 24) & 0x7fu);
        const int pr = static_cast<int>((meta >> 16) & 0xffu);
        const double w = static_cast<double>(B.rules[meta & 0xffu].weight);
#pragma unroll
        for (int v = 0; v < BRM_MAX_REWARD; ++v)
          if (v == pr) r[v] = __dadd_rn(r[v], w);
      }
      if (nxt != cur) s.q[i] = (s.q[i] & ~(15ull << (4 * sl))) | (static_cast<uint64_t>(nxt) << (4 * sl));
    }
    // GetActionCost (:120-132); explicit _rn ops: no FMA contraction, so the double
    // result is bit-identical to the reference's separate multiply and add.
    {
      const int d = delta[i];
      const int dd = d - last_in[i];
      const double cost = __dadd_rn(__dmul_rn(-static_cast<double>(abs(d - 1)), B.w_speed),
                                    __dmul_rn(-static_cast<double>(dd * dd), B.w_acc));
      const int vlast = B.V - 1;
#pragma unroll
      for (int v = 0; v < BRM_MAX_REWARD; ++v)
        if (v == vlast) r[v] = __dadd_rn(r[v], cost);
    }
    if (coll || depth_hit) term_out |= 1u << i;  // :82-84
    sink(i, r);
  }
  s.term = term_out;
  s.depth += 1;
}

// GetTerminalReward (:141-152) of agent i
template <int A>
__device__ __forceinline__ void gw_terminal_reward(const GwBlob& B, const GwRegs<A>& s, int i,
                                                   double (&r)[BRM_MAX_REWARD]) {
#pragma unroll
  for (int v = 0; v < BRM_MAX_REWARD; ++v) r[v] = 0.0;
  for (int sl = 0; sl < B.R; ++sl) {
    const brm_rule& rule = B.rules[B.slot_rule[i][sl]];
    const uint32_t cur = static_cast<uint32_t>(s.q[i] >> (4 * sl)) & 15u;
    const double pen = rule.accepting[cur] ? 0.0 : static_cast<double>(rule.weight);
    const int pr = rule.priority;
#pragma unroll
    for (int v = 0; v < BRM_MAX_REWARD; ++v)
      if (v == pr) r[v] = __dadd_rn(r[v], pen);
  }
}

}  // namespace brm
