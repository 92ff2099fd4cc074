// Continuous-world host API (include/brm.h, brm_cw_*): scenario blobs, staging and the
// batched MvmctsState surface.  Kernels live in cw_kernels.cu.
#include <algorithm>
#include <cmath>
#include <cstddef>

#include "common.cuh"
#include "cw_device.cuh"
#include "staging.cuh"

namespace brm {

int cw_launch_rollout(brm_engine* e, const CwLaunchCtx& ctx, int n_scenarios, const brm_cw_states& leaves,
                      const brm_rollout_params& rp, const brm_cw_rollout_out& out);
int cw_launch_batch(brm_engine* e, const CwLaunchCtx& ctx, int mode, const brm_cw_states& in,
                    const brm_cw_states* out, const uint8_t* actions, float* rewards,
                    uint8_t* num_actions, float* term_reward, int do_check_terminal, int32_t T,
                    const brm_cw_trace* tr);
struct CwEngine {
  int n_scenarios = 0;
  uint8_t* d_blobs = nullptr;
  int32_t* d_blob_bytes = nullptr;
  CwLaunchCtx ctx;
};

namespace {

constexpr int kThreads = kCwThreads;

bool rule_specific(const brm_rule& r) {
  for (int a = 0; a < r.n_aps; ++a)
    if (r.ap_agent_specific[a]) return true;
  return false;
}

int build_blob(const brm_cw_scenario& s, CwBlob* B, int* R_out) {
  const brm_cw_params& p = s.params;
  memset(B, 0, offsetof(CwBlob, segs));
  BRM_REQUIRE(s.n_agents >= 1 && s.n_agents <= BRM_CW_MAX_AGENTS, "brm_cw_configure: n_agents %d not in [1,%d]",
              s.n_agents, BRM_CW_MAX_AGENTS);
  BRM_REQUIRE(s.n_mcts_agents >= 1 && s.n_mcts_agents <= s.n_agents,
              "brm_cw_configure: n_mcts_agents %d not in [1,n_agents]", s.n_mcts_agents);
  BRM_REQUIRE(p.reward_vector_size >= 1 && p.reward_vector_size <= BRM_MAX_REWARD,
              "brm_cw_configure: reward_vector_size %d not in [1,%d]", p.reward_vector_size, BRM_MAX_REWARD);
  BRM_REQUIRE(s.n_lanes >= 0 && s.n_lanes <= BRM_CW_MAX_LANES, "brm_cw_configure: bad n_lanes %d", s.n_lanes);
  BRM_REQUIRE(s.n_road >= 3 && s.n_road <= BRM_CW_MAX_ROAD, "brm_cw_configure: road polygon needs 3..%d vertices",
              BRM_CW_MAX_ROAD);
  BRM_REQUIRE(s.n_rules >= 0 && s.n_rules <= BRM_MAX_RULES, "brm_cw_configure: bad n_rules %d", s.n_rules);
  BRM_REQUIRE(p.prediction_time_span > 0.f && p.integration_time_delta > 0.f && p.wheel_base > 0.f,
              "brm_cw_configure: time spans and wheel base must be positive");
  B->A = s.n_agents;
  B->M = s.n_mcts_agents;
  B->V = p.reward_vector_size;
  B->n_lanes = s.n_lanes;
  B->n_road = s.n_road;
  B->n_rules = s.n_rules;
  B->ego_only = p.use_rule_reward_for_ego_only ? 1 : 0;
  // BehaviorMPContinuousActions::Plan: ceil(dt / integration_dt) Euler sub-steps, the last
  // one takes the remainder (bark, un-vendored)
  const double dt = p.prediction_time_span, hs = p.integration_time_delta;
  int n_sub = static_cast<int>(std::ceil(dt / hs - 1e-9));
  if (n_sub < 1) n_sub = 1;
  BRM_REQUIRE(n_sub <= 1024, "brm_cw_configure: %d integration sub-steps per step is not sensible", n_sub);
  B->n_sub = n_sub;
  B->h_sub = static_cast<float>(hs);
  B->h_last = static_cast<float>(dt - (n_sub - 1) * hs);
  {
    // closed form of the sub-stepped Euler scheme for primitives without steering
    // (cw_integrate): T = sum h_n, C2 = sum h_n t_n, with the float-rounded step sizes the
    // kernel would otherwise use
    double T = 0.0, C2 = 0.0;
    for (int n = 0; n < n_sub; ++n) {
      const double h = (n == n_sub - 1) ? static_cast<double>(B->h_last) : static_cast<double>(B->h_sub);
      C2 += h * T;
      T += h;
    }
    B->t_total = static_cast<float>(T);
    B->c2 = static_cast<float>(C2);
  }
  B->dt = p.prediction_time_span;
  B->inv_dt = 1.0f / p.prediction_time_span;
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
  BRM_REQUIRE(p.sd_max_decel_rear != 0.f && p.sd_max_decel_front != 0.f,
              "brm_cw_configure: safe-distance decelerations must be non-zero");
  B->sd_ar = static_cast<float>(1.0 / (2.0 * std::fabs(static_cast<double>(p.sd_max_decel_rear))));
  B->sd_af = static_cast<float>(1.0 / (2.0 * std::fabs(static_cast<double>(p.sd_max_decel_front))));
  B->near2 = p.near_distance * p.near_distance;
  for (int i = 0; i < s.n_agents; ++i) {
    const brm_cw_agent& ag = s.agents[i];
    BRM_REQUIRE(ag.n_prims >= 1 && ag.n_prims <= BRM_CW_MAX_PRIMS, "brm_cw_configure: agent %d has %d primitives",
                i, ag.n_prims);
    B->front[i] = ag.front;
    B->rear[i] = ag.rear;
    B->hw[i] = ag.half_width;
    {
      const double ex = 0.5 * (static_cast<double>(ag.front) + ag.rear), ey = ag.half_width;
      B->rad[i] = static_cast<float>(std::sqrt(ex * ex + ey * ey) * (1.0 + 1e-5) + 1e-4);
    }
    B->has_goal[i] = ag.has_goal ? 1 : 0;
    for (int k = 0; k < 6; ++k) B->goal[i][k] = ag.goal[k];
    B->n_prims[i] = static_cast<uint8_t>(ag.n_prims);
    for (int k = 0; k < ag.n_prims; ++k) {
      B->prim_a[i][k] = ag.prim_acc[k];
      B->prim_k[i][k] = static_cast<float>(std::tan(static_cast<double>(ag.prim_delta[k])) /
                                           static_cast<double>(p.wheel_base));
    }
  }
  // passive agents (not in the MCTS agent list) are dealt round-robin to the M lanes of a rollout
  {
    const int M = s.n_mcts_agents, P = s.n_agents - M;
    B->Ps = (P + M - 1) / M;
    for (int j = M; j < s.n_agents; ++j) {
      B->pas_owner[j] = static_cast<uint8_t>((j - M) % M);
      B->pas_slot[j] = static_cast<uint8_t>((j - M) / M);
    }
  }
  // rule slots of an agent: rules in order, agent-specific ones once per other world agent
  // ascending (BehaviorRulesMcts::MakeRuleStates, src/behavior_rules_mcts.hpp:295-322)
  int R = 0;
  for (int k = 0; k < s.n_rules; ++k) {
    const brm_rule& r = s.rules[k];
    BRM_REQUIRE(r.n_aps <= BRM_MAX_APS && r.n_states >= 1 && r.n_states <= BRM_MAX_DFA_STATES,
                "brm_cw_configure: rule %d has bad sizes", k);
    BRM_REQUIRE(r.priority < p.reward_vector_size, "brm_cw_configure: rule %d priority %d >= V", k, r.priority);
    BRM_REQUIRE(r.init_state < r.n_states, "brm_cw_configure: rule %d init_state out of range", k);
    for (int a = 0; a < r.n_aps; ++a) {
      const int kind = r.ap_label[a];
      const bool ego_kind = kind <= BRM_CW_LABEL_ON_ROAD;
      const bool other_kind = kind >= BRM_CW_LABEL_MERGED_X && kind <= BRM_CW_LABEL_NEAR_X;
      BRM_REQUIRE(ego_kind || other_kind, "brm_cw_configure: rule %d AP %d: unknown label kind %d", k, a, kind);
      BRM_REQUIRE(other_kind == (r.ap_agent_specific[a] != 0),
                  "brm_cw_configure: rule %d AP %d: label kind %d %s a #0 placeholder", k, a, kind,
                  other_kind ? "needs" : "cannot take");
    }
    for (int t = 0; t < r.n_states * (1 << r.n_aps); ++t)
      BRM_REQUIRE(r.trans[t] == BRM_VIOLATION || r.trans[t] < r.n_states,
                  "brm_cw_configure: rule %d transition %d out of range", k, t);
    B->rules[k] = r;
    // device order of the letter bits: AP k at bit k (the ABI's tables have AP 0 at the top bit)
    for (int q = 0; q < r.n_states; ++q)
      for (int l = 0; l < (1 << r.n_aps); ++l) {
        int nl = 0;
        for (int a = 0; a < r.n_aps; ++a) nl |= ((l >> (r.n_aps - 1 - a)) & 1) << a;
        B->rules[k].trans[q * (1 << r.n_aps) + nl] = r.trans[q * (1 << r.n_aps) + l];
      }
    const int reps = rule_specific(r) ? s.n_agents - 1 : 1;
    for (int j = 0; j < reps; ++j) {
      BRM_REQUIRE(R < BRM_CW_MAX_SLOTS, "brm_cw_configure: more than %d rule states per agent", BRM_CW_MAX_SLOTS);
      B->slot_rule[R] = static_cast<uint8_t>(k);
      B->slot_rank[R] = rule_specific(r) ? static_cast<int8_t>(j) : -1;
      uint4 rec;
      // x: per AP the bit of the slot's label word it reads; unused APs read bit 31 (always 0)
      rec.x = 0;
      for (int a = 0; a < BRM_MAX_APS; ++a)
        rec.x |= static_cast<uint32_t>(a < r.n_aps ? r.ap_label[a] : 31) << (8 * a);
      rec.y = static_cast<uint32_t>(r.n_aps) | (static_cast<uint32_t>(r.priority) << 8) |
              (static_cast<uint32_t>(r.init_state) << 16) | (static_cast<uint32_t>(r.on_violation ? 1 : 0) << 24);
      rec.z = static_cast<uint32_t>(k) | (static_cast<uint32_t>(rule_specific(r) ? j + 1 : 0) << 8);
      // bits 16..31: states that loop to themselves on every letter (absorbing, no violation)
      for (int q = 0; q < r.n_states; ++q) {
        bool sink = true;
        for (int l = 0; l < (1 << r.n_aps); ++l) sink = sink && r.trans[q * (1 << r.n_aps) + l] == q;
        if (sink) rec.z |= 1u << (16 + q);
      }
      memcpy(&rec.w, &r.weight, sizeof(float));
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
    B->road_edge[e] = make_float4(static_cast<float>(ya), static_cast<float>(yb), static_cast<float>(xa),
                                  yb != ya ? static_cast<float>((xb - xa) / (yb - ya)) : 0.f);
  }
  // lane centre lines -> per-segment records
  int off = 0, n_blk = 0;
  for (int l = 0; l < s.n_lanes; ++l) {
    const brm_cw_lane& L = s.lanes[l];
    BRM_REQUIRE(L.n_pts >= 2 && L.n_pts <= BRM_CW_MAX_PTS, "brm_cw_configure: lane %d has %d points", l, L.n_pts);
    BRM_REQUIRE(L.successor >= -1 && L.successor < s.n_lanes, "brm_cw_configure: lane %d: bad successor", l);
    CwLaneHdr& H = B->lanes[l];
    H.seg_off = off;
    H.n_seg = L.n_pts - 1;
    H.succ = L.successor;
    H.flags = L.flags;
    H.hw = L.half_width;
    H.hw2 = L.half_width * L.half_width;
    H.s_point = L.s_point;
    double s0 = 0.0;
    for (int k = 0; k + 1 < L.n_pts; ++k) {
      const double x0 = L.pts[k][0], y0 = L.pts[k][1], x1 = L.pts[k + 1][0], y1 = L.pts[k + 1][1];
      const double dx = x1 - x0, dy = y1 - y0, l2 = dx * dx + dy * dy;
      BRM_REQUIRE(l2 > 0.0, "brm_cw_configure: lane %d has a zero-length segment at point %d", l, k);
      CwSeg& g = B->segs[off++];
      g.x0 = static_cast<float>(x0);
      g.y0 = static_cast<float>(y0);
      g.dx = static_cast<float>(dx);
      g.dy = static_cast<float>(dy);
      g.inv_len2 = static_cast<float>(1.0 / l2);
      g.len = static_cast<float>(std::sqrt(l2));
      g.s0 = static_cast<float>(s0);
      g._pad = 0.f;
      s0 += std::sqrt(l2);
    }
    H.length = static_cast<float>(s0);
    // bounding boxes for the exact pruning in cw_frenet (grown a little: a larger box only
    // prunes less, never differently)
    const double kSlack = 1e-3;
    const double grow = L.half_width + 2.0 * kSlack;
    for (int b0 = 0; b0 < L.n_pts - 1; b0 += kCwBlk) {
      const int b1 = std::min(b0 + kCwBlk, L.n_pts - 1);
      double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
      for (int k = b0; k <= b1; ++k) {
        x0 = std::min(x0, static_cast<double>(L.pts[k][0]));
        x1 = std::max(x1, static_cast<double>(L.pts[k][0]));
        y0 = std::min(y0, static_cast<double>(L.pts[k][1]));
        y1 = std::max(y1, static_cast<double>(L.pts[k][1]));
      }
      BRM_REQUIRE(n_blk < kCwMaxBlkTotal, "brm_cw_configure: too many lane segments (more than %d blocks)",
                  kCwMaxBlkTotal);
      B->blk_box[n_blk] = make_float4(static_cast<float>(x0 - kSlack), static_cast<float>(y0 - kSlack),
                                      static_cast<float>(x1 + kSlack), static_cast<float>(y1 + kSlack));
      B->blk_gbox[n_blk] = make_float4(static_cast<float>(x0 - grow), static_cast<float>(y0 - grow),
                                       static_cast<float>(x1 + grow), static_cast<float>(y1 + grow));
      B->blk_info[n_blk] = static_cast<uint32_t>(l) | (static_cast<uint32_t>(b0) << 8) |
                           (static_cast<uint32_t>(b1 - b0) << 16);
      ++n_blk;
    }
  }
  B->n_blk = n_blk;
  // grid over the union of the grown boxes (a little larger, so that a point on or beyond its
  // border is outside every box); a cell lists the blocks whose grown box overlaps the cell
  // widened by a margin that covers the fp32 evaluation of the cell index on the device
  {
    double gx0 = 1e300, gy0 = 1e300, gx1 = -1e300, gy1 = -1e300;
    for (int i = 0; i < n_blk; ++i) {
      gx0 = std::min(gx0, static_cast<double>(B->blk_gbox[i].x));
      gy0 = std::min(gy0, static_cast<double>(B->blk_gbox[i].y));
      gx1 = std::max(gx1, static_cast<double>(B->blk_gbox[i].z));
      gy1 = std::max(gy1, static_cast<double>(B->blk_gbox[i].w));
    }
    if (n_blk == 0) gx0 = gy0 = 0.0, gx1 = gy1 = 1.0;
    gx0 -= 0.01; gy0 -= 0.01; gx1 += 0.01; gy1 += 0.01;
    const double cw = (gx1 - gx0) / 8.0, ch = (gy1 - gy0) / 8.0;
    B->grid_x0 = static_cast<float>(gx0);
    B->grid_y0 = static_cast<float>(gy0);
    B->grid_ix = static_cast<float>(1.0 / cw);
    B->grid_iy = static_cast<float>(1.0 / ch);
    // the device's cell of a coordinate: floor((x - fl(gx0)) * fl(1/cw)) in fp32; its error is far
    // below one percent of a cell
    const double ox = B->grid_x0, oy = B->grid_y0, sx = 1.0 / B->grid_ix, sy = 1.0 / B->grid_iy;
    int most = 0;
    for (int cy = 0; cy < 8; ++cy)
      for (int cx = 0; cx < 8; ++cx) {
        const double x0 = ox + (cx - 0.01) * sx, x1 = ox + (cx + 1.01) * sx;
        const double y0 = oy + (cy - 0.01) * sy, y1 = oy + (cy + 1.01) * sy;
        unsigned long long m = 0ull;
        for (int i = 0; i < n_blk; ++i) {
          const float4 g = B->blk_gbox[i];
          if (g.x <= x1 && g.z >= x0 && g.y <= y1 && g.w >= y0) m |= 1ull << i;
        }
        B->cell_mask[cy * 8 + cx] = m;
        most = std::max(most, __builtin_popcountll(m));
      }
    B->use_grid = (n_blk > 0 && 4 * most <= n_blk) ? 1 : 0;
  }
  B->total_bytes = static_cast<int32_t>(offsetof(CwBlob, segs) + static_cast<size_t>(off) * sizeof(CwSeg));
  return BRM_OK;
}

CwEngine* cw_of(brm_engine* e) { return static_cast<CwEngine*>(e->cw); }

int check_ready(const brm_engine* e, const char* fn) {
  if (!e) {
    set_error("%s: engine is NULL", fn);
    return BRM_ERR_INVALID;
  }
  if (!e->cw) {
    set_error("%s: call brm_cw_configure first", fn);
    return BRM_ERR_NOT_CONFIGURED;
  }
  return BRM_OK;
}

int check_states(const brm_engine* e, const brm_cw_states* s, const char* fn, bool homogeneous) {
  BRM_REQUIRE(s != nullptr, "%s: states is NULL", fn);
  BRM_REQUIRE(s->n > 0, "%s: empty batch (n=%d)", fn, s->n);
  BRM_REQUIRE(s->scenario && s->agents && s->flags && s->horizon && s->terminal && s->q,
              "%s: scenario/agents/flags/horizon/terminal/q must be set", fn);
  BRM_REQUIRE(s->mem == BRM_MEM_HOST || s->mem == BRM_MEM_DEVICE, "%s: bad mem %d", fn, s->mem);
  if (s->mem == BRM_MEM_HOST) {
    const int ns = static_cast<CwEngine*>(e->cw)->n_scenarios;
    for (int i = 0; i < s->n; ++i) {
      BRM_REQUIRE(s->scenario[i] >= 0 && s->scenario[i] < ns, "%s: state %d: scenario %d out of range [0,%d)", fn,
                  i, s->scenario[i], ns);
      if (homogeneous)
        BRM_REQUIRE(s->scenario[i] == s->scenario[0],
                    "%s: all states of one batch must live in the same scenario (state %d)", fn, i);
    }
  }
  return BRM_OK;
}

int upload_states(Stager& st, const brm_cw_states& h, int A, int M, int R, brm_cw_states* d) {
  const size_t n = h.n;
  *d = h;
  d->mem = BRM_MEM_DEVICE;
  BRM_TRY(st.in(h.scenario, n, &d->scenario));
  BRM_TRY(st.in(h.agents, A * n * 4, &d->agents));
  BRM_TRY(st.in(h.flags, A * n, &d->flags));
  BRM_TRY(st.in(h.horizon, n, &d->horizon));
  BRM_TRY(st.in(h.terminal, n, &d->terminal));
  BRM_TRY(st.in(h.q, static_cast<size_t>(M) * R * n, &d->q));
  if (h.viol) BRM_TRY(st.in(h.viol, static_cast<size_t>(M) * R * n, &d->viol));
  return BRM_OK;
}

int alloc_states(Stager& st, const brm_cw_states& like, int A, int M, int R, brm_cw_states* d) {
  const size_t n = like.n;
  *d = like;
  d->mem = BRM_MEM_DEVICE;
  BRM_TRY(st.alloc(n, &d->scenario));
  BRM_TRY(st.alloc(A * n * 4, &d->agents));
  BRM_TRY(st.alloc(A * n, &d->flags));
  BRM_TRY(st.alloc(n, &d->horizon));
  BRM_TRY(st.alloc(n, &d->terminal));
  BRM_TRY(st.alloc(static_cast<size_t>(M) * R * n, &d->q));
  if (like.viol) BRM_TRY(st.alloc(static_cast<size_t>(M) * R * n, &d->viol, true));
  return BRM_OK;
}

int download_states(Stager& st, const brm_cw_states& d, int A, int M, int R, brm_cw_states* h) {
  const size_t n = h->n;
  BRM_TRY(st.out(h->scenario, d.scenario, n));
  BRM_TRY(st.out(h->agents, d.agents, A * n * 4));
  BRM_TRY(st.out(h->flags, d.flags, A * n));
  BRM_TRY(st.out(h->horizon, d.horizon, n));
  BRM_TRY(st.out(h->terminal, d.terminal, n));
  BRM_TRY(st.out(h->q, d.q, static_cast<size_t>(M) * R * n));
  if (h->viol) BRM_TRY(st.out(h->viol, d.viol, static_cast<size_t>(M) * R * n));
  return BRM_OK;
}

}  // namespace
}  // namespace brm

using namespace brm;

extern "C" {

void brm_cw_release(brm_engine* e) {
  if (!e || !e->cw) return;
  CwEngine* c = cw_of(e);
  if (c->d_blobs) cudaFree(c->d_blobs);
  if (c->d_blob_bytes) cudaFree(c->d_blob_bytes);
  delete c;
  e->cw = nullptr;
}

int brm_cw_configure(brm_engine* e, const brm_cw_scenario* scenarios, int32_t n_scenarios) {
  BRM_REQUIRE(e && scenarios && n_scenarios >= 1, "brm_cw_configure: bad arguments");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  std::vector<CwBlob>* blobs = new std::vector<CwBlob>(n_scenarios);
  std::vector<int32_t> bytes(n_scenarios);
  int A = 0, M = 0, V = 0, R = 0, cap = 0, use_grid = 1;
  for (int i = 0; i < n_scenarios; ++i) {
    int Ri = 0;
    int rc = build_blob(scenarios[i], &(*blobs)[i], &Ri);
    if (rc != BRM_OK) {
      delete blobs;
      return rc;
    }
    const CwBlob& B = (*blobs)[i];
    if (i == 0) {
      A = B.A; M = B.M; V = B.V; R = Ri;
    } else if (B.A != A || B.M != M || B.V != V || Ri != R) {
      delete blobs;
      set_error("brm_cw_configure: scenario %d differs from scenario 0 in n_agents / n_mcts_agents / "
                "reward_vector_size / rule states per agent", i);
      return BRM_ERR_INVALID;
    }
    bytes[i] = (B.total_bytes + 15) & ~15;
    if (bytes[i] > cap) cap = bytes[i];
    use_grid &= B.use_grid;
  }
  brm_cw_release(e);
  CwEngine* c = new CwEngine();
  c->n_scenarios = n_scenarios;
  const size_t stride = sizeof(CwBlob);
  cudaError_t err = cudaMalloc(&c->d_blobs, stride * n_scenarios);
  if (err == cudaSuccess) err = cudaMalloc(&c->d_blob_bytes, sizeof(int32_t) * n_scenarios);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(c->d_blobs, blobs->data(), stride * n_scenarios, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(c->d_blob_bytes, bytes.data(), sizeof(int32_t) * n_scenarios, cudaMemcpyHostToDevice,
                          e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  delete blobs;
  if (err != cudaSuccess) {
    set_error("brm_cw_configure: %s", cudaGetErrorString(err));
    if (c->d_blobs) cudaFree(c->d_blobs);
    if (c->d_blob_bytes) cudaFree(c->d_blob_bytes);
    delete c;
    return BRM_ERR_CUDA;
  }
  c->ctx.c.blobs = c->d_blobs;
  c->ctx.c.blob_bytes = c->d_blob_bytes;
  c->ctx.c.blob_stride = stride;
  c->ctx.c.blob_cap = cap;
  c->ctx.c.A = A;
  c->ctx.c.M = M;
  c->ctx.c.V = V;
  c->ctx.c.R = R;
  c->ctx.c.G = 32 / M;  // one lane per MCTS agent
  c->ctx.c.Ps = (A - M + M - 1) / M;
  c->ctx.c.use_grid = use_grid;
  c->ctx.smem_bytes = static_cast<size_t>(cap) + 2u * static_cast<size_t>(R) * kThreads +
                      static_cast<size_t>(c->ctx.c.Ps) * kCwPasWords * sizeof(float) * kThreads;
  e->cw = c;
  return BRM_OK;
}

int32_t brm_cw_rules_per_agent(const brm_engine* e) {
  return (e && e->cw) ? static_cast<CwEngine*>(e->cw)->ctx.c.R : -1;
}

int brm_cw_check_terminal(brm_engine* e, brm_cw_states* s) {
  BRM_TRY(check_ready(e, "brm_cw_check_terminal"));
  BRM_TRY(check_states(e, s, "brm_cw_check_terminal", true));
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.c.A, M = c->ctx.c.M, R = c->ctx.c.R;
  if (s->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_QUERY, *s, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, nullptr);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *s, A, M, R, &d));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_QUERY, d, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, nullptr));
  BRM_TRY(st.out(s->terminal, d.terminal, static_cast<size_t>(s->n)));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_execute(brm_engine* e, const brm_cw_states* in, const uint8_t* actions, brm_cw_states* out,
                   float* rewards) {
  BRM_TRY(check_ready(e, "brm_cw_execute"));
  BRM_TRY(check_states(e, in, "brm_cw_execute", true));
  BRM_REQUIRE(out && out->scenario && out->agents && out->flags && out->horizon && out->terminal && out->q,
              "brm_cw_execute: every array of `out` except viol is required");
  BRM_REQUIRE(actions && rewards, "brm_cw_execute: actions/rewards is NULL");
  BRM_REQUIRE(in->n == out->n && in->mem == out->mem, "brm_cw_execute: in/out batch mismatch");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.c.A, M = c->ctx.c.M, R = c->ctx.c.R, V = c->ctx.c.V;
  const size_t n = in->n;
  if (in->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_EXECUTE, *in, out, actions, rewards, nullptr, nullptr, 0, 0, nullptr);
  for (size_t i = 0; i < n; ++i)
    if (in->terminal[i]) {  // BARK_EXPECT_TRUE(!IsTerminal()), src/rules_mcts_state.cpp:59
      set_error("brm_cw_execute: state %zu is terminal", i);
      return BRM_ERR_TERMINAL;
    }
  Stager st(e);
  brm_cw_states d_in, d_out;
  uint8_t* d_act;
  float* d_rew;
  BRM_TRY(upload_states(st, *in, A, M, R, &d_in));
  BRM_TRY(alloc_states(st, *out, A, M, R, &d_out));
  BRM_TRY(st.in(actions, M * n, &d_act));
  BRM_TRY(st.alloc(M * V * n, &d_rew, true));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_EXECUTE, d_in, &d_out, d_act, d_rew, nullptr, nullptr, 0, 0, nullptr));
  BRM_TRY(download_states(st, d_out, A, M, R, out));
  BRM_TRY(st.out(rewards, d_rew, M * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_evaluate_rules(brm_engine* e, const brm_cw_states* in, brm_cw_states* out, float* rewards) {
  BRM_TRY(check_ready(e, "brm_cw_evaluate_rules"));
  BRM_TRY(check_states(e, in, "brm_cw_evaluate_rules", true));
  BRM_REQUIRE(out && out->scenario && out->agents && out->flags && out->horizon && out->terminal && out->q,
              "brm_cw_evaluate_rules: every array of `out` except viol is required");
  BRM_REQUIRE(rewards, "brm_cw_evaluate_rules: rewards is NULL");
  BRM_REQUIRE(in->n == out->n && in->mem == out->mem, "brm_cw_evaluate_rules: in/out batch mismatch");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.c.A, M = c->ctx.c.M, R = c->ctx.c.R, V = c->ctx.c.V;
  const size_t n = in->n;
  if (in->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_RULES, *in, out, nullptr, rewards, nullptr, nullptr, 0, 0, nullptr);
  Stager st(e);
  brm_cw_states d_in, d_out;
  float* d_rew;
  BRM_TRY(upload_states(st, *in, A, M, R, &d_in));
  BRM_TRY(alloc_states(st, *out, A, M, R, &d_out));
  BRM_TRY(st.alloc(M * V * n, &d_rew, true));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_RULES, d_in, &d_out, nullptr, d_rew, nullptr, nullptr, 0, 0, nullptr));
  BRM_TRY(download_states(st, d_out, A, M, R, out));
  BRM_TRY(st.out(rewards, d_rew, M * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

static int cw_query(brm_engine* e, const brm_cw_states* s, uint8_t* num_actions, float* term_reward,
                    const char* fn) {
  BRM_TRY(check_ready(e, fn));
  BRM_TRY(check_states(e, s, fn, true));
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.c.A, M = c->ctx.c.M, R = c->ctx.c.R, V = c->ctx.c.V;
  const size_t n = s->n;
  if (s->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_QUERY, *s, nullptr, nullptr, nullptr, num_actions, term_reward, 0, 0,
                           nullptr);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *s, A, M, R, &d));
  uint8_t* d_na = nullptr;
  float* d_tr = nullptr;
  if (num_actions) BRM_TRY(st.alloc(M * n, &d_na));
  if (term_reward) BRM_TRY(st.alloc(M * V * n, &d_tr));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_QUERY, d, nullptr, nullptr, nullptr, d_na, d_tr, 0, 0, nullptr));
  if (num_actions) BRM_TRY(st.out(num_actions, d_na, M * n));
  if (term_reward) BRM_TRY(st.out(term_reward, d_tr, M * V * n));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_get_num_actions(brm_engine* e, const brm_cw_states* s, uint8_t* out) {
  BRM_REQUIRE(out, "brm_cw_get_num_actions: out is NULL");
  return cw_query(e, s, out, nullptr, "brm_cw_get_num_actions");
}

int brm_cw_terminal_reward(brm_engine* e, const brm_cw_states* s, float* out) {
  BRM_REQUIRE(out, "brm_cw_terminal_reward: out is NULL");
  return cw_query(e, s, nullptr, out, "brm_cw_terminal_reward");
}

// rollouts, optionally followed by the root statistics (first_action != NULL) in the same call
static int cw_rollout_impl(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                           const brm_cw_rollout_out* out, const uint8_t* first_action, int32_t n_actions,
                           double* stats) {
  BRM_TRY(check_ready(e, "brm_cw_rollout"));
  BRM_TRY(check_states(e, leaves, "brm_cw_rollout", false));
  BRM_REQUIRE(rp && out && (out->returns_sum || first_action), "brm_cw_rollout: rp/out/returns_sum is required");
  BRM_REQUIRE(rp->rollouts_per_leaf >= 1, "brm_cw_rollout: rollouts_per_leaf must be >= 1");
  BRM_REQUIRE(rp->max_steps >= 0 && rp->max_steps <= 255, "brm_cw_rollout: max_steps %d not in [0,255]",
              rp->max_steps);
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.c.A, M = c->ctx.c.M, R = c->ctx.c.R, V = c->ctx.c.V;
  if (leaves->mem == BRM_MEM_DEVICE) {
    BRM_REQUIRE(out->returns_sum, "brm_cw_rollout: device batches need a returns_sum buffer");
    BRM_TRY(cw_launch_rollout(e, c->ctx, c->n_scenarios, *leaves, *rp, *out));
    if (first_action)
      BRM_TRY(brm::launch_root_stats(e, leaves->n, M, n_actions, V, first_action, out->returns_sum,
                                     rp->rollouts_per_leaf, stats));
    return BRM_OK;
  }
  const size_t L = leaves->n, N = L * static_cast<size_t>(rp->rollouts_per_leaf);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *leaves, A, M, R, &d));
  brm_cw_rollout_out o = {};
  BRM_TRY(st.alloc(L * M * V, &o.returns_sum));
  if (out->viol_sum) BRM_TRY(st.alloc(L * M * R, &o.viol_sum));
  if (out->agent_steps) BRM_TRY(st.alloc(L, &o.agent_steps));
  if (out->ro_returns) BRM_TRY(st.alloc(N * M * V, &o.ro_returns));
  if (out->ro_viol) BRM_TRY(st.alloc(N * M * R, &o.ro_viol));
  if (out->ro_q) BRM_TRY(st.alloc(N * M * R, &o.ro_q));
  if (out->ro_agents) BRM_TRY(st.alloc(N * A * 4, &o.ro_agents));
  if (out->ro_flags) BRM_TRY(st.alloc(N * A, &o.ro_flags));
  if (out->ro_steps) BRM_TRY(st.alloc(N, &o.ro_steps));
  BRM_TRY(cw_launch_rollout(e, c->ctx, c->n_scenarios, d, *rp, o));
  if (first_action) {
    uint8_t* d_fa;
    double* d_st;
    const size_t st_count = static_cast<size_t>(M) * n_actions * (1 + V);
    BRM_TRY(st.in(first_action, L * M, &d_fa));
    BRM_TRY(st.in(stats, st_count, &d_st));
    BRM_TRY(brm::launch_root_stats(e, leaves->n, M, n_actions, V, d_fa, o.returns_sum, rp->rollouts_per_leaf, d_st));
    BRM_TRY(st.out(stats, d_st, st_count));
  }
  if (out->returns_sum) BRM_TRY(st.out(out->returns_sum, o.returns_sum, L * M * V));
  if (out->viol_sum) BRM_TRY(st.out(out->viol_sum, o.viol_sum, L * M * R));
  if (out->agent_steps) BRM_TRY(st.out(out->agent_steps, o.agent_steps, L));
  if (out->ro_returns) BRM_TRY(st.out(out->ro_returns, o.ro_returns, N * M * V));
  if (out->ro_viol) BRM_TRY(st.out(out->ro_viol, o.ro_viol, N * M * R));
  if (out->ro_q) BRM_TRY(st.out(out->ro_q, o.ro_q, N * M * R));
  if (out->ro_agents) BRM_TRY(st.out(out->ro_agents, o.ro_agents, N * A * 4));
  if (out->ro_flags) BRM_TRY(st.out(out->ro_flags, o.ro_flags, N * A));
  if (out->ro_steps) BRM_TRY(st.out(out->ro_steps, o.ro_steps, N));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

int brm_cw_rollout(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                   const brm_cw_rollout_out* out) {
  return cw_rollout_impl(e, leaves, rp, out, nullptr, 0, nullptr);
}

int brm_cw_rollout_root_stats(brm_engine* e, const brm_cw_states* leaves, const brm_rollout_params* rp,
                              const brm_cw_rollout_out* out, const uint8_t* leaf_first_action, int32_t n_actions,
                              double* stats) {
  BRM_REQUIRE(leaf_first_action && stats && n_actions >= 1, "brm_cw_rollout_root_stats: bad arguments");
  return cw_rollout_impl(e, leaves, rp, out, leaf_first_action, n_actions, stats);
}

int brm_cw_replay(brm_engine* e, const brm_cw_states* s0, const uint8_t* actions, int32_t T,
                  const brm_cw_trace* tr) {
  BRM_TRY(check_ready(e, "brm_cw_replay"));
  BRM_TRY(check_states(e, s0, "brm_cw_replay", true));
  BRM_REQUIRE(actions && tr && T >= 1, "brm_cw_replay: bad arguments");
  BRM_REQUIRE(tr->agents_out && tr->flags_out && tr->terminal_out && tr->q_out && tr->viol_out &&
                  tr->rewards_out && tr->labels_out && tr->frenet_out && tr->term_reward_out && tr->steps_out,
              "brm_cw_replay: every trace array is required");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  CwEngine* c = cw_of(e);
  const int A = c->ctx.c.A, M = c->ctx.c.M, R = c->ctx.c.R, V = c->ctx.c.V;
  if (s0->mem == BRM_MEM_DEVICE)
    return cw_launch_batch(e, c->ctx, MODE_REPLAY, *s0, nullptr, actions, nullptr, nullptr, nullptr, 0, T, tr);
  const size_t B = s0->n, BT = B * static_cast<size_t>(T);
  Stager st(e);
  brm_cw_states d;
  BRM_TRY(upload_states(st, *s0, A, M, R, &d));
  uint8_t* d_act;
  BRM_TRY(st.in(actions, BT * M, &d_act));
  brm_cw_trace o = {};
  BRM_TRY(st.alloc(BT * A * 4, &o.agents_out, true));
  BRM_TRY(st.alloc(BT * A, &o.flags_out, true));
  BRM_TRY(st.alloc(BT, &o.terminal_out, true));
  BRM_TRY(st.alloc(BT * M * R, &o.q_out, true));
  BRM_TRY(st.alloc(BT * M * R, &o.viol_out, true));
  BRM_TRY(st.alloc(BT * M * V, &o.rewards_out, true));
  BRM_TRY(st.alloc(BT * M * 4, &o.labels_out, true));
  BRM_TRY(st.alloc(BT * A * 3, &o.frenet_out, true));
  BRM_TRY(st.alloc(B * M * V, &o.term_reward_out, true));
  BRM_TRY(st.alloc(B, &o.steps_out, true));
  BRM_TRY(cw_launch_batch(e, c->ctx, MODE_REPLAY, d, nullptr, d_act, nullptr, nullptr, nullptr, 0, T, &o));
  BRM_TRY(st.out(tr->agents_out, o.agents_out, BT * A * 4));
  BRM_TRY(st.out(tr->flags_out, o.flags_out, BT * A));
  BRM_TRY(st.out(tr->terminal_out, o.terminal_out, BT));
  BRM_TRY(st.out(tr->q_out, o.q_out, BT * M * R));
  BRM_TRY(st.out(tr->viol_out, o.viol_out, BT * M * R));
  BRM_TRY(st.out(tr->rewards_out, o.rewards_out, BT * M * V));
  BRM_TRY(st.out(tr->labels_out, o.labels_out, BT * M * 4));
  BRM_TRY(st.out(tr->frenet_out, o.frenet_out, BT * A * 3));
  BRM_TRY(st.out(tr->term_reward_out, o.term_reward_out, B * M * V));
  BRM_TRY(st.out(tr->steps_out, o.steps_out, B));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

}  // extern "C"
