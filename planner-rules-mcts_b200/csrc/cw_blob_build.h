// Host-side construction of the scenario blob (CwBlob, cw_core.cuh) from the ABI's
// brm_cw_scenario.  Plain C++ (no CUDA types): used by cw_engine.cu and by the CPU harness of
// the tests (tests/cpp/cw_emul.cpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>

#include "cw_core.cuh"

namespace brm {

inline bool cw_rule_specific(const brm_rule& r) {
  for (int a = 0; a < r.n_aps; ++a)
    if (r.ap_agent_specific[a]) return true;
  return false;
}

#define BRM_BLOB_REQUIRE(cond, ...)                    \
  do {                                                 \
    if (!(cond)) {                                     \
      char _buf[384];                                  \
      snprintf(_buf, sizeof(_buf), __VA_ARGS__);       \
      *err = _buf;                                     \
      return false;                                    \
    }                                                  \
  } while (0)

// Fills *B (and the number of rule states per agent); on failure returns false with a message.
inline bool cw_build_blob(const brm_cw_scenario& s, CwBlob* B, int* R_out, std::string* err) {
  const brm_cw_params& p = s.params;
  memset(B, 0, offsetof(CwBlob, segs));
  BRM_BLOB_REQUIRE(s.n_agents >= 1 && s.n_agents <= BRM_CW_MAX_AGENTS, "brm_cw_configure: n_agents %d not in [1,%d]",
                   s.n_agents, BRM_CW_MAX_AGENTS);
  BRM_BLOB_REQUIRE(s.n_mcts_agents >= 1 && s.n_mcts_agents <= s.n_agents,
                   "brm_cw_configure: n_mcts_agents %d not in [1,n_agents]", s.n_mcts_agents);
  BRM_BLOB_REQUIRE(p.reward_vector_size >= 1 && p.reward_vector_size <= BRM_MAX_REWARD,
                   "brm_cw_configure: reward_vector_size %d not in [1,%d]", p.reward_vector_size, BRM_MAX_REWARD);
  BRM_BLOB_REQUIRE(s.n_lanes >= 0 && s.n_lanes <= BRM_CW_MAX_LANES, "brm_cw_configure: bad n_lanes %d", s.n_lanes);
  BRM_BLOB_REQUIRE(s.n_road >= 3 && s.n_road <= BRM_CW_MAX_ROAD,
                   "brm_cw_configure: road polygon needs 3..%d vertices", BRM_CW_MAX_ROAD);
  BRM_BLOB_REQUIRE(s.n_rules >= 0 && s.n_rules <= BRM_MAX_RULES, "brm_cw_configure: bad n_rules %d", s.n_rules);
  BRM_BLOB_REQUIRE(p.prediction_time_span > 0.0 && p.integration_time_delta > 0.0 && p.wheel_base > 0.0,
                   "brm_cw_configure: time spans and wheel base must be positive");
  B->A = s.n_agents;
  B->M = s.n_mcts_agents;
  B->V = p.reward_vector_size;
  B->n_lanes = s.n_lanes;
  B->n_road = s.n_road;
  B->n_rules = s.n_rules;
  B->ego_only = p.use_rule_reward_for_ego_only ? 1 : 0;
  B->remove_oom = p.remove_agents_out_of_map ? 1 : 0;
  // BehaviorMPContinuousActions::Plan: ceil(dt / integration_dt) Euler sub-steps, the last
  // one takes the remainder (bark, un-vendored)
  const double dt = p.prediction_time_span, hs = p.integration_time_delta;
  int n_sub = static_cast<int>(std::ceil(dt / hs - 1e-9));
  if (n_sub < 1) n_sub = 1;
  BRM_BLOB_REQUIRE(n_sub <= 1024, "brm_cw_configure: %d integration sub-steps per step is not sensible", n_sub);
  B->n_sub = n_sub;
  B->h_sub = hs;
  B->h_last = dt - (n_sub - 1) * hs;
  {
    // closed form of the sub-stepped Euler scheme (cw_integrate): T = sum h_n, C2 = sum h_n t_n
    double T = 0.0, C2 = 0.0;
    for (int n = 0; n < n_sub; ++n) {
      const double h = (n == n_sub - 1) ? B->h_last : B->h_sub;
      C2 += h * T;
      T += h;
    }
    B->t_total = T;
    B->c2 = C2;
  }
  B->dt = dt;
  B->inv_dt = 1.0 / dt;
  B->w_coll = p.collision_weight;
  B->w_oom = p.out_of_map_weight;
  B->w_pot = p.potential_weight;
  B->w_acc = p.acceleration_weight;
  B->w_rad = p.radial_acceleration_weight;
  B->w_vel = p.desired_velocity_weight;
  B->w_lane = p.lane_center_weight;
  B->v_des = p.desired_velocity;
  B->gamma = p.discount_factor;
  B->goal_reward = p.goal_reward;
  B->sd_delta = p.sd_reaction_time;
  BRM_BLOB_REQUIRE(p.sd_max_decel_rear != 0.0 && p.sd_max_decel_front != 0.0,
                   "brm_cw_configure: safe-distance decelerations must be non-zero");
  B->sd_ar = 1.0 / (2.0 * std::fabs(p.sd_max_decel_rear));
  B->sd_af = 1.0 / (2.0 * std::fabs(p.sd_max_decel_front));
  B->near2 = p.near_distance * p.near_distance;
  auto coord_ok = [](double v) { return std::fabs(v) <= BRM_CW_MAX_COORD; };
  for (int i = 0; i < s.n_agents; ++i) {
    const brm_cw_agent& ag = s.agents[i];
    BRM_BLOB_REQUIRE(ag.n_prims >= 1 && ag.n_prims <= BRM_CW_MAX_PRIMS,
                     "brm_cw_configure: agent %d has %d primitives", i, ag.n_prims);
    B->front[i] = ag.front;
    B->rear[i] = ag.rear;
    B->hw[i] = ag.half_width;
    B->off[i] = (ag.front - ag.rear) * 0.5;
    B->ext[i] = (ag.front + ag.rear) * 0.5;
    B->rad[i] = std::sqrt(B->ext[i] * B->ext[i] + ag.half_width * ag.half_width) * (1.0 + 1e-9) + 1e-9;
    B->has_goal[i] = ag.has_goal ? 1 : 0;
    for (int k = 0; k < 6; ++k) B->goal[i][k] = ag.goal[k];
    B->n_prims[i] = static_cast<uint8_t>(ag.n_prims);
    for (int k = 0; k < ag.n_prims; ++k) {
      B->prim_a[i][k] = ag.prim_acc[k];
      B->prim_k[i][k] = std::tan(ag.prim_delta[k]) / p.wheel_base;
    }
  }
  // rule slots of an agent: rules in order, agent-specific ones once per other world agent
  // ascending (BehaviorRulesMcts::MakeRuleStates, src/behavior_rules_mcts.hpp:295-322)
  int R = 0;
  for (int k = 0; k < s.n_rules; ++k) {
    const brm_rule& r = s.rules[k];
    BRM_BLOB_REQUIRE(r.n_aps <= BRM_MAX_APS && r.n_states >= 1 && r.n_states <= BRM_MAX_DFA_STATES,
                     "brm_cw_configure: rule %d has bad sizes", k);
    BRM_BLOB_REQUIRE(r.priority < p.reward_vector_size, "brm_cw_configure: rule %d priority %d >= V", k, r.priority);
    BRM_BLOB_REQUIRE(r.init_state < r.n_states, "brm_cw_configure: rule %d init_state out of range", k);
    for (int a = 0; a < r.n_aps; ++a) {
      const int kind = r.ap_label[a];
      const bool ego_kind = kind <= BRM_CW_LABEL_ON_ROAD;
      const bool other_kind = kind >= BRM_CW_LABEL_MERGED_X && kind <= BRM_CW_LABEL_NEAR_X;
      BRM_BLOB_REQUIRE(ego_kind || other_kind, "brm_cw_configure: rule %d AP %d: unknown label kind %d", k, a, kind);
      BRM_BLOB_REQUIRE(other_kind == (r.ap_agent_specific[a] != 0),
                       "brm_cw_configure: rule %d AP %d: label kind %d %s a #0 placeholder", k, a, kind,
                       other_kind ? "needs" : "cannot take");
    }
    for (int t = 0; t < r.n_states * (1 << r.n_aps); ++t)
      BRM_BLOB_REQUIRE(r.trans[t] == BRM_VIOLATION || r.trans[t] < r.n_states,
                       "brm_cw_configure: rule %d transition %d out of range", k, t);
    B->rules[k] = r;
    // device order of the letter bits: AP k at bit k (the ABI's tables have AP 0 at the top bit)
    for (int q = 0; q < r.n_states; ++q)
      for (int l = 0; l < (1 << r.n_aps); ++l) {
        int nl = 0;
        for (int a = 0; a < r.n_aps; ++a) nl |= ((l >> (r.n_aps - 1 - a)) & 1) << a;
        B->rules[k].trans[q * (1 << r.n_aps) + nl] = r.trans[q * (1 << r.n_aps) + l];
      }
    const bool specific = cw_rule_specific(r);
    const int reps = specific ? s.n_agents - 1 : 1;
    for (int j = 0; j < reps; ++j) {
      BRM_BLOB_REQUIRE(R < BRM_CW_MAX_SLOTS, "brm_cw_configure: more than %d rule states per agent", BRM_CW_MAX_SLOTS);
      B->slot_rule[R] = static_cast<uint8_t>(k);
      CwSlotRec rec;
      rec.x = 0;
      for (int a = 0; a < BRM_MAX_APS; ++a)
        rec.x |= static_cast<uint32_t>(a < r.n_aps ? r.ap_label[a] : 31) << (8 * a);
      rec.y = static_cast<uint32_t>(r.n_aps) | (static_cast<uint32_t>(r.priority) << 8) |
              (static_cast<uint32_t>(r.init_state) << 16) | (static_cast<uint32_t>(r.on_violation ? 1 : 0) << 24);
      rec.z = static_cast<uint32_t>(k) | (static_cast<uint32_t>(specific ? j + 1 : 0) << 8);
      // bits 16..31: states that loop to themselves on every letter (absorbing, no violation)
      for (int q = 0; q < r.n_states; ++q) {
        bool sink = true;
        for (int l = 0; l < (1 << r.n_aps); ++l) sink = sink && r.trans[q * (1 << r.n_aps) + l] == q;
        if (sink) rec.z |= 1u << (16 + q);
      }
      rec.w = r.weight;
      B->slot_rec[R] = rec;
      ++R;
    }
  }
  B->R = R;
  *R_out = R;
  // road polygon edges (ya, yb, xa, m)
  for (int e = 0; e < s.n_road; ++e) {
    const int f = (e + 1) % s.n_road;
    const double xa = s.road[e][0], ya = s.road[e][1], xb = s.road[f][0], yb = s.road[f][1];
    BRM_BLOB_REQUIRE(coord_ok(xa) && coord_ok(ya), "brm_cw_configure: road vertex %d outside +-%g m", e,
                     BRM_CW_MAX_COORD);
    CwEdge g;
    g.xa = xa;
    g.ya = ya;
    g.yb = yb;
    g.m = (yb != ya) ? (xb - xa) / (yb - ya) : 0.0;
    B->road_edge[e] = g;
  }
  // lane centre lines -> per-segment fields (stored field by field, n_seg_total apart)
  int ns_total = 0;
  for (int l = 0; l < s.n_lanes; ++l) ns_total += std::max(0, s.lanes[l].n_pts - 1);
  BRM_BLOB_REQUIRE(ns_total <= kCwMaxSegs, "brm_cw_configure: %d lane segments (limit %d)", ns_total, kCwMaxSegs);
  B->n_seg_total = ns_total;
  auto seg = [&](int field, int k) -> double& { return B->segs[static_cast<size_t>(field) * ns_total + k]; };
  int off = 0;
  for (int l = 0; l < s.n_lanes; ++l) {
    const brm_cw_lane& L = s.lanes[l];
    BRM_BLOB_REQUIRE(L.n_pts >= 2 && L.n_pts <= BRM_CW_MAX_PTS, "brm_cw_configure: lane %d has %d points", l, L.n_pts);
    BRM_BLOB_REQUIRE(L.successor >= -1 && L.successor < s.n_lanes, "brm_cw_configure: lane %d: bad successor", l);
    BRM_BLOB_REQUIRE(L.half_width >= 0.0 && L.half_width <= 64.0, "brm_cw_configure: lane %d: half width %g", l,
                     L.half_width);
    CwLaneHdr& H = B->lanes[l];
    H.seg_off = off;
    H.n_seg = L.n_pts - 1;
    H.succ = L.successor;
    H.flags = L.flags;
    H.hw2 = L.half_width * L.half_width;
    H.s_point = L.s_point;
    double s0 = 0.0;
    for (int k = 0; k < L.n_pts; ++k)
      BRM_BLOB_REQUIRE(coord_ok(L.pts[k][0]) && coord_ok(L.pts[k][1]),
                       "brm_cw_configure: lane %d point %d outside +-%g m", l, k, BRM_CW_MAX_COORD);
    for (int k = 0; k + 1 < L.n_pts; ++k) {
      const double x0 = L.pts[k][0], y0 = L.pts[k][1], x1 = L.pts[k + 1][0], y1 = L.pts[k + 1][1];
      const int g = off++;
      const double dx = x1 - x0, dy = y1 - y0;
      const double l2 = dx * dx + dy * dy;
      BRM_BLOB_REQUIRE(l2 > 0.0, "brm_cw_configure: lane %d has a zero-length segment at point %d", l, k);
      const double len = std::sqrt(l2);
      seg(SEG_X0, g) = x0;
      seg(SEG_Y0, g) = y0;
      seg(SEG_DX, g) = dx;
      seg(SEG_DY, g) = dy;
      seg(SEG_INV, g) = 1.0 / l2;
      seg(SEG_LEN, g) = len;
      seg(SEG_S0, g) = s0;
      s0 += len;
    }
    H.length = s0;
    // cw_frenet: grown bounding box of the lane and the segment ranges of its buckets.  A point
    // is tested as its fp32 rounding (error below 2.5e-4 m up to BRM_CW_MAX_COORD) with fp32
    // arithmetic (a few 1e-4 m more), so everything is grown by kSlack: growing only skips
    // less, never differently.
    const double kSlack = 4e-3;
    const double grow = L.half_width + kSlack;
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    for (int k = 0; k < L.n_pts; ++k) {
      x0 = std::min(x0, L.pts[k][0]);
      x1 = std::max(x1, L.pts[k][0]);
      y0 = std::min(y0, L.pts[k][1]);
      y1 = std::max(y1, L.pts[k][1]);
    }
    H.gx0 = static_cast<float>(x0 - grow);
    H.gy0 = static_cast<float>(y0 - grow);
    H.gx1 = static_cast<float>(x1 + grow);
    H.gy1 = static_cast<float>(y1 + grow);
    const int axis = (x1 - x0) >= (y1 - y0) ? 0 : 1;
    const double c_lo = (axis ? y0 : x0) - grow - kSlack, c_hi = (axis ? y1 : x1) + grow + kSlack;
    const double width = (c_hi - c_lo) / kCwBuckets;
    H.b_axis = axis;
    H.b_org = static_cast<float>(c_lo);
    H.b_inv = static_cast<float>(1.0 / width);
    int lo[kCwBuckets], hi[kCwBuckets];
    for (int b = 0; b < kCwBuckets; ++b) lo[b] = 255, hi[b] = -1;
    for (int k = 0; k + 1 < L.n_pts; ++k) {
      const double a0 = std::min(L.pts[k][axis], L.pts[k + 1][axis]) - grow;
      const double a1 = std::max(L.pts[k][axis], L.pts[k + 1][axis]) + grow;
      const int b0 = std::max(0, static_cast<int>(std::floor((a0 - c_lo) / width)));
      const int b1 = std::min(kCwBuckets - 1, static_cast<int>(std::floor((a1 - c_lo) / width)));
      for (int b = b0; b <= b1; ++b) {
        lo[b] = std::min(lo[b], k);
        hi[b] = std::max(hi[b], k);
      }
    }
    for (int b = 0; b < kCwBuckets; ++b)
      B->bucket[l][b] = hi[b] < 0 ? static_cast<uint16_t>(1) /* lo 1 > hi 0: none */
                                  : static_cast<uint16_t>(lo[b] | (hi[b] << 8));
  }
  // cw_classify_shape: grid over the road polygon.  A cell is INSIDE / OUTSIDE when no edge of
  // the polygon comes within kCellSlack of it (so that every point the fp32 cell index can map
  // into the cell is decided like the cell's centre); VERTEX when a road vertex lies within
  // the largest agent's reach of the cell (only then can a vertex be inside a rectangle whose
  // centre is in the cell).
  {
    const double kCellSlack = 8e-3;
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    for (int e = 0; e < s.n_road; ++e) {
      x0 = std::min(x0, s.road[e][0]);
      x1 = std::max(x1, s.road[e][0]);
      y0 = std::min(y0, s.road[e][1]);
      y1 = std::max(y1, s.road[e][1]);
    }
    const double margin = 0.05;
    x0 -= margin; y0 -= margin; x1 += margin; y1 += margin;
    const double cw = (x1 - x0) / kCwGrid, ch = (y1 - y0) / kCwGrid;
    B->grid_x0 = static_cast<float>(x0);
    B->grid_y0 = static_cast<float>(y0);
    B->grid_ix = static_cast<float>(1.0 / cw);
    B->grid_iy = static_cast<float>(1.0 / ch);
    double reach = 0.0;
    for (int i = 0; i < s.n_agents; ++i) reach = std::max(reach, std::fabs(B->off[i]) + B->rad[i]);
    auto seg_box_dist = [](double ax, double ay, double bx, double by, double rx0, double ry0, double rx1,
                           double ry1) {
      // distance between the segment a-b and the rectangle, by sampling-free clipping: 0 if the
      // segment enters the rectangle, else the smallest corner / end point distance
      auto pt_rect = [&](double px, double py) {
        const double dx = std::max(std::max(rx0 - px, px - rx1), 0.0), dy = std::max(std::max(ry0 - py, py - ry1), 0.0);
        return std::sqrt(dx * dx + dy * dy);
      };
      auto pt_seg = [&](double px, double py) {
        const double vx = bx - ax, vy = by - ay, l2 = vx * vx + vy * vy;
        double t = l2 > 0.0 ? ((px - ax) * vx + (py - ay) * vy) / l2 : 0.0;
        t = std::min(std::max(t, 0.0), 1.0);
        const double ex = px - (ax + t * vx), ey = py - (ay + t * vy);
        return std::sqrt(ex * ex + ey * ey);
      };
      // Liang-Barsky: does the segment pass through the rectangle?
      double t0 = 0.0, t1 = 1.0;
      const double dx = bx - ax, dy = by - ay;
      const double pp[4] = {-dx, dx, -dy, dy}, qq[4] = {ax - rx0, rx1 - ax, ay - ry0, ry1 - ay};
      bool hit = true;
      for (int i = 0; i < 4 && hit; ++i) {
        if (pp[i] == 0.0) {
          if (qq[i] < 0.0) hit = false;
        } else {
          const double r = qq[i] / pp[i];
          if (pp[i] < 0.0) t0 = std::max(t0, r); else t1 = std::min(t1, r);
          if (t0 > t1) hit = false;
        }
      }
      if (hit) return 0.0;
      double d = std::min(pt_rect(ax, ay), pt_rect(bx, by));
      d = std::min(d, std::min(std::min(pt_seg(rx0, ry0), pt_seg(rx1, ry0)), std::min(pt_seg(rx0, ry1), pt_seg(rx1, ry1))));
      return d;
    };
    auto inside = [&](double px, double py) {
      bool in = false;
      for (int e = 0; e < s.n_road; ++e) {
        const int f = (e + 1) % s.n_road;
        const double xa = s.road[e][0], ya = s.road[e][1], xb = s.road[f][0], yb = s.road[f][1];
        if ((ya > py) != (yb > py) && px < (xb - xa) / (yb - ya) * (py - ya) + xa) in = !in;
      }
      return in;
    };
    memset(B->grid, 0, sizeof(B->grid));
    for (int gy = 0; gy < kCwGrid; ++gy)
      for (int gx = 0; gx < kCwGrid; ++gx) {
        const double rx0 = x0 + gx * cw, rx1 = rx0 + cw, ry0 = y0 + gy * ch, ry1 = ry0 + ch;
        double d_edge = 1e300, d_vert = 1e300;
        for (int e = 0; e < s.n_road; ++e) {
          const int f = (e + 1) % s.n_road;
          d_edge = std::min(d_edge, seg_box_dist(s.road[e][0], s.road[e][1], s.road[f][0], s.road[f][1], rx0, ry0,
                                                 rx1, ry1));
          const double vx = std::max(std::max(rx0 - s.road[e][0], s.road[e][0] - rx1), 0.0);
          const double vy = std::max(std::max(ry0 - s.road[e][1], s.road[e][1] - ry1), 0.0);
          d_vert = std::min(d_vert, std::sqrt(vx * vx + vy * vy));
        }
        uint32_t code = 0;
        if (d_edge > kCellSlack) code |= inside(0.5 * (rx0 + rx1), 0.5 * (ry0 + ry1)) ? CW_CELL_INSIDE : CW_CELL_OUTSIDE;
        if (d_vert <= reach + kCellSlack) code |= CW_CELL_VERTEX;
        const int c = gy * kCwGrid + gx;
        B->grid[c >> 3] |= code << ((c & 7) * 4);
      }
  }
  B->total_bytes = static_cast<int32_t>(offsetof(CwBlob, segs) +
                                       static_cast<size_t>(SEG_FIELDS) * ns_total * sizeof(double));
  return true;
}

}  // namespace brm
