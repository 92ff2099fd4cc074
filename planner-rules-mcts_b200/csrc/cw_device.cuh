// Device-side continuous-world step: MvmctsState::Execute
// (/root/reference/src/rules_mcts_state.cpp:57-128) with one lane = one MCTS agent of one
// rollout.  floor(32/M) rollouts share a warp (M = agents in the reference's agent_idx
// list); the agents that are only simulated (constant-acceleration prediction, not in
// agent_idx) are dealt round-robin to the lanes of their rollout, which keep them in
// shared-memory columns and step them in a loop.  Agents of a rollout exchange poses with
// warp shuffles, pairwise predicates (collision, preceding agent, near) are evaluated by
// every lane against its group, the rule automata of an agent step in its lane.
// The scenario blob (parameters, lane polylines, road polygon, automaton tables) sits in
// shared memory, staged by a TMA bulk copy.
#pragma once
#include "common.cuh"

namespace brm {

struct alignas(16) CwSeg {
  float x0, y0, dx, dy;
  float inv_len2, len, s0, _pad;
};

struct alignas(16) CwLaneHdr {
  int32_t seg_off, n_seg, succ, flags;
  float hw, hw2, length, s_point;
};

constexpr int kCwMaxSegs = BRM_CW_MAX_LANES * (BRM_CW_MAX_PTS - 1);
constexpr int kCwThreads = 128;          // CTA size of every continuous-world kernel
// resident CTAs per SM the rollout kernel is compiled for.  Measured on B200 (M / W / S
// workloads): 5 CTAs (<= 102 registers, 96 used) beats 6 (80 regs) and 8 (64 regs) by 2-12 %,
// and 4 or 3 CTAs are no better -- registers matter more than occupancy here.
#ifndef BRM_CW_CTAS_PER_SM
#define BRM_CW_CTAS_PER_SM 5
#endif
constexpr int kCwRolloutCtasPerSm = BRM_CW_CTAS_PER_SM;
constexpr int kCwBlk = 8;                                        // segments per pruning block
constexpr int kCwMaxBlk = (BRM_CW_MAX_PTS - 1 + kCwBlk - 1) / kCwBlk;  // blocks per lane
constexpr int kCwMaxBlkTotal = BRM_CW_MAX_LANES * kCwMaxBlk;            // 64: one bit each

struct alignas(16) CwBlob {
  int32_t A, M, V, R, n_lanes, n_road, n_rules, n_sub;
  int32_t ego_only, total_bytes, n_blk, Ps;  // Ps: passive-agent slots per lane = ceil((A-M)/M)
  float w_coll, w_oom, w_pot, w_acc, w_rad, w_vel, w_lane, dt;
  float v_des, gamma, goal_reward, h_sub, h_last, sd_delta, sd_ar, sd_af;
  float near2, inv_dt, _p2, _p3;
  float front[BRM_CW_MAX_AGENTS], rear[BRM_CW_MAX_AGENTS], hw[BRM_CW_MAX_AGENTS];
  float rad[BRM_CW_MAX_AGENTS];  // circumscribed radius of the rectangle (a hair more), about its centre
  float goal[BRM_CW_MAX_AGENTS][6];
  uint8_t has_goal[BRM_CW_MAX_AGENTS], n_prims[BRM_CW_MAX_AGENTS];
  // passive agent j >= M lives in slot pas_slot[j] of the lane with agent index pas_owner[j]
  uint8_t pas_owner[BRM_CW_MAX_AGENTS], pas_slot[BRM_CW_MAX_AGENTS];
  float prim_a[BRM_CW_MAX_AGENTS][BRM_CW_MAX_PRIMS];
  float prim_k[BRM_CW_MAX_AGENTS][BRM_CW_MAX_PRIMS];  // tan(delta) / wheel_base
  uint8_t slot_rule[BRM_CW_MAX_SLOTS];
  int8_t slot_rank[BRM_CW_MAX_SLOTS];  // rank of the bound other agent among the others, -1: ego rule
  brm_rule rules[BRM_MAX_RULES];
  uint4 slot_rec[BRM_CW_MAX_SLOTS];  // compact per-slot record read by cw_evaluate
  float c2, t_total, _p4, _p5;       // closed form of the Euler sub-steps without steering
  float4 road_edge[BRM_CW_MAX_ROAD];  // (ya, yb, xa, m = (xb-xa)/(yb-ya))
  // exact pruning of the nearest-segment search: (xmin, ymin, xmax, ymax) of every lane
  // (grown by its half width) and of every block of kCwBlk consecutive segments
  // blocks of all lanes in lane order (n_blk <= 64, one bit each in the candidate mask):
  // blk_gbox = bounding box grown by the lane's half width, blk_box = tight box,
  // blk_info = lane | first segment (within the lane) << 8 | segment count << 16
  // an 8 x 8 grid over the union of the grown boxes: per cell the blocks whose grown box reaches
  // into it (a superset; the blocks named there are then tested exactly)
  // (`use_grid`, decided at upload: used when the longest list of a cell is at most a quarter of
  // the blocks; where many lanes cross, testing all boxes in a uniform loop is cheaper than a
  // divergent walk over the cell's list -- the rollout kernel has an instance for either)
  float grid_x0, grid_y0, grid_ix, grid_iy;  // origin and 1 / cell size
  int32_t use_grid, _p6, _p7, _p8;
  unsigned long long cell_mask[64];
  float4 blk_gbox[kCwMaxBlkTotal];
  float4 blk_box[kCwMaxBlkTotal];
  uint32_t blk_info[kCwMaxBlkTotal];
  CwLaneHdr lanes[BRM_CW_MAX_LANES];
  CwSeg segs[kCwMaxSegs];  // last: only the used part is copied
};
static_assert(sizeof(CwBlob) % 16 == 0, "blob must be a multiple of 16 bytes");
static_assert(offsetof(CwBlob, segs) % 16 == 0, "segment array must be 16-byte aligned");

// launch context shared by cw_engine.cu (host API) and cw_kernels.cu
struct CwCommon {
  const uint8_t* blobs;       // [n_scenarios] CwBlob, stride blob_stride
  const int32_t* blob_bytes;  // used bytes of each blob (multiple of 16)
  size_t blob_stride;
  int32_t blob_cap;           // bytes reserved for the blob in shared memory
  int32_t A, M, V, R, G;      // G = rollouts (states) per warp = 32 / M
  int32_t Ps;                 // passive-agent slots per lane
  int32_t use_grid;           // every scenario's lane projection takes its candidates from the grid
};
struct CwLaunchCtx {
  CwCommon c;
  size_t smem_bytes;
};
enum { MODE_EXECUTE = 0, MODE_QUERY = 1, MODE_REPLAY = 2, MODE_RULES = 3 };

// per-lane agent state
struct CwAgent {
  float x, y, th, v;
  float cs, sn;     // cos / sin of th
  uint32_t flags;   // BRM_CW_PRESENT | BRM_CW_VALID
  int lane;         // lane index of the current position (-1: none)
  float s, lat;
  float ex, ey, eth, ev;  // Kahan compensation of x, y, th, v (carried across steps)
};

struct CwStepOut {
  float r[BRM_MAX_REWARD];
  uint32_t w[4];  // labels: ego bits by kind, masks over j for kinds 8, 9, 10
  bool terminal;  // state-level (group-uniform)
};

// compensated (Kahan) accumulation: sum += term with the rounding error carried in `c`.
// The reference integrates in double; fp32 states that simply accumulate 15 sub-steps per
// step drift by ~1e-4 m over a rollout, which is visible in rewards that difference two
// nearly equal speeds.  With the compensation the fp32 state stays within ~1 ulp of the
// exactly accumulated value.
__device__ __forceinline__ void kahan_add(float& sum, float& c, float term) {
  const float y = __fsub_rn(term, c);
  const float t = __fadd_rn(sum, y);
  c = __fsub_rn(__fsub_rn(t, sum), y);
  sum = t;
}

// reward vectors are four named scalars (an array indexed by a runtime priority would be
// placed in local memory): component `idx` += val
struct CwVec {
  float c0, c1, c2, c3;
};
__device__ __forceinline__ void vec_add_at(CwVec& r, int idx, float val) {
  r.c0 = idx == 0 ? r.c0 + val : r.c0;
  r.c1 = idx == 1 ? r.c1 + val : r.c1;
  r.c2 = idx == 2 ? r.c2 + val : r.c2;
  r.c3 = idx == 3 ? r.c3 + val : r.c3;
}

__device__ __forceinline__ void cw_integrate(const CwBlob& B, CwAgent& t, int a, uint32_t prim) {
  const float acc = B.prim_a[a][prim];
  const float k = B.prim_k[a][prim];
  float x = t.x, y = t.y, th = t.th, v = t.v;
  float ex = t.ex, ey = t.ey, eth = t.eth, ev = t.ev;
  float cs = t.cs, sn = t.sn;
  const int n_sub = B.n_sub;
  if (k == 0.0f) {
    // no steering: theta never changes, so the explicit-Euler sub-steps telescope exactly:
    //   v_n = v0 + a t_n,  x_N = x0 + cos(theta) * sum_n h_n v_n = x0 + cos(theta) (v0 T + a C2)
    // with T = sum_n h_n and C2 = sum_n h_n t_n (t_n = time at the start of sub-step n), both
    // computed on the host in double.  Same scheme as the reference's loop, one rounding
    // instead of fifteen.
    // explicit _rn intrinsics: the compiler may not contract them differently in different
    // kernels, so rollout, replay and single-step kernels produce the same bits
    const float d = __fmaf_rn(v, B.t_total, __fmul_rn(acc, B.c2));
    kahan_add(x, ex, __fmul_rn(cs, d));
    kahan_add(y, ey, __fmul_rn(sn, d));
    kahan_add(v, ev, __fmul_rn(acc, B.t_total));
  } else {
    // steering: the speed and the heading at the start of sub-step n are closed forms of the
    // same telescoping sums, v_n = v0 + a t_n and theta_n = theta0 + k (v0 t_n + a c2_n) with
    // c2_n = sum_{m<n} h_m t_m, so only the displacement sum_n h_n v_n (cos, sin)(theta_n) is
    // accumulated in the loop (15 terms of ~0.2 m) and every state component is updated with ONE
    // compensated addition per step.  Inside the step the heading only feeds one 0.02 s
    // displacement, so the SFU sine / cosine (abs. error 2^-21.4 on [-pi, pi]) is enough; the
    // pose keeps the accurate sincosf below.
    float dx = 0.f, dy = 0.f, tn = 0.f, c2n = 0.f;
    for (int n = 0; n < n_sub; ++n) {
      const float h = (n == n_sub - 1) ? B.h_last : B.h_sub;
      const float vn = __fmaf_rn(acc, tn, v);
      if (n > 0) {
        const float thn = __fmaf_rn(k, __fmaf_rn(v, tn, __fmul_rn(acc, c2n)), th);
        if (fabsf(thn) <= 3.2f)
          __sincosf(thn, &sn, &cs);
        else
          sincosf(thn, &sn, &cs);
      }
      const float hv = __fmul_rn(h, vn);
      dx = __fmaf_rn(hv, cs, dx);
      dy = __fmaf_rn(hv, sn, dy);
      c2n = __fmaf_rn(h, tn, c2n);
      tn = __fadd_rn(tn, h);
    }
    const float d = __fmaf_rn(v, B.t_total, __fmul_rn(acc, B.c2));
    kahan_add(x, ex, dx);
    kahan_add(y, ey, dy);
    kahan_add(th, eth, __fmul_rn(k, d));
    kahan_add(v, ev, __fmul_rn(acc, B.t_total));
    sincosf(th, &sn, &cs);
  }
  t.x = x; t.y = y; t.th = th; t.v = v; t.cs = cs; t.sn = sn;
  t.ex = ex; t.ey = ey; t.eth = eth; t.ev = ev;
}

// Nearest point on every lane centre line; the first lane that contains the point wins.
// The reference semantics is a brute-force nearest-segment search per lane (that is what
// the oracle does).  Two exact prunings keep the same answer:
//   * only blocks of segments whose bounding box, grown by the lane's half width, contains
//     the point can hold a segment within the half width (candidate mask, built with
//     uniform control flow over all blocks of the scenario);
//   * a candidate block whose tight box is farther than the best distance found so far on
//     its lane cannot hold the nearest segment.
// A lane none of whose blocks survives is a lane the point is not on.
template <bool GRID = false>
__device__ __forceinline__ void cw_frenet(const CwBlob& B, CwAgent& t) {
  t.lane = -1;
  t.s = 0.f;
  t.lat = 0.f;
  // candidate mask: the grid cell of the point names the blocks that can contain it, those are
  // tested exactly (a point outside the grid is outside every grown box)
  unsigned long long mask = 0ull;
  if (GRID) {
    const float fx = (t.x - B.grid_x0) * B.grid_ix, fy = (t.y - B.grid_y0) * B.grid_iy;
    unsigned long long cand = 0ull;
    if (fx >= 0.f && fx < 8.f && fy >= 0.f && fy < 8.f)
      cand = B.cell_mask[static_cast<int>(fy) * 8 + static_cast<int>(fx)];
    while (cand) {
      const int i = __ffsll(static_cast<long long>(cand)) - 1;
      cand &= cand - 1ull;
      const float4 g = B.blk_gbox[i];
      if (t.x >= g.x && t.x <= g.z && t.y >= g.y && t.y <= g.w) mask |= 1ull << i;
    }
  } else {
    // all boxes, in two 32-bit halves (scenarios rarely have more than 32 blocks)
    const int nb = B.n_blk;
    uint32_t lo = 0u, hi = 0u, bit = 1u;
    const int n0 = nb < 32 ? nb : 32;
    for (int i = 0; i < n0; ++i, bit <<= 1) {
      const float4 g = B.blk_gbox[i];
      if (t.x >= g.x && t.x <= g.z && t.y >= g.y && t.y <= g.w) lo |= bit;
    }
    bit = 1u;
    for (int i = 32; i < nb; ++i, bit <<= 1) {
      const float4 g = B.blk_gbox[i];
      if (t.x >= g.x && t.x <= g.z && t.y >= g.y && t.y <= g.w) hi |= bit;
    }
    mask = static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
  }
  const float kInf = __int_as_float(0x7f800000);
  int l = -1, n_seg = 0, seg_off = 0, bk = 0;
  float best = kInf, btu = 0.f, hw2 = 0.f;
  for (;;) {
    const int i = mask ? (__ffsll(static_cast<long long>(mask)) - 1) : -1;
    const uint32_t info = i >= 0 ? B.blk_info[i] : 0xffu;
    const int li = static_cast<int>(info & 0xffu);
    if (li != l) {
      // lane l is finished: is the point on it?
      if (l >= 0 && best <= hw2) {
        const bool interior = !(bk == 0 && btu < 0.f) && !(bk == n_seg - 1 && btu > 1.f);
        if (interior) {
          const CwSeg g = B.segs[seg_off + bk];
          const float bt = fminf(fmaxf(btu, 0.f), 1.f);
          const float wx = t.x - g.x0, wy = t.y - g.y0;
          const float cross = g.dx * wy - g.dy * wx;
          t.lane = l;
          t.s = g.s0 + bt * g.len;
          const float d = sqrtf(best);
          t.lat = cross < 0.f ? -d : d;
          return;
        }
      }
      if (i < 0) return;
      l = li;
      const CwLaneHdr L = B.lanes[l];
      n_seg = L.n_seg;
      seg_off = L.seg_off;
      hw2 = L.hw2;
      best = kInf;
    }
    mask &= mask - 1ull;
    const float4 bb = B.blk_box[i];
    const float gx = fmaxf(fmaxf(bb.x - t.x, t.x - bb.z), 0.f);
    const float gy = fmaxf(fmaxf(bb.y - t.y, t.y - bb.w), 0.f);
    if (gx * gx + gy * gy > best) continue;
    const int b0 = static_cast<int>((info >> 8) & 0xffu);
    const int b1 = b0 + static_cast<int>(info >> 16);
    const CwSeg* seg = B.segs + seg_off;
    for (int k = b0; k < b1; ++k) {
      const float4 p = reinterpret_cast<const float4*>(seg + k)[0];
      const float inv = seg[k].inv_len2;
      const float wx = t.x - p.x, wy = t.y - p.y;
      const float tu = (wx * p.z + wy * p.w) * inv;
      const float tc = fminf(fmaxf(tu, 0.f), 1.f);
      const float ex = wx - tc * p.z, ey = wy - tc * p.w;
      const float d2 = ex * ex + ey * ey;
      if (d2 < best) {
        best = d2;
        bk = k;
        btu = tu;
      }
    }
  }
}

__device__ __forceinline__ bool cw_point_in_road(const CwBlob& B, float px, float py) {
  bool inside = false;
  for (int e = 0; e < B.n_road; ++e) {
    const float4 g = B.road_edge[e];
    if ((g.x > py) != (g.y > py)) {
      if (px < g.w * (py - g.x) + g.z) inside = !inside;
    }
  }
  return inside;
}

__device__ __forceinline__ void cw_corners(const CwBlob& B, const CwAgent& t, int a, float (&cx)[4],
                                           float (&cy)[4]) {
  const float f = B.front[a], r = B.rear[a], h = B.hw[a];
  const float lx[4] = {f, f, -r, -r};
  const float ly[4] = {h, -h, -h, h};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cx[k] = t.x + (lx[k] * t.cs - ly[k] * t.sn);
    cy[k] = t.y + (lx[k] * t.sn + ly[k] * t.cs);
  }
}

// EvaluatorDrivableArea restated: a corner outside the road polygon, or a road vertex
// strictly inside the rectangle
__device__ __forceinline__ bool cw_out_of_map(const CwBlob& B, const CwAgent& t, int a,
                                              const float (&cx)[4], const float (&cy)[4]) {
  // one pass over the road edges: crossing-number parity of the four corners (the same
  // test as cw_point_in_road) and "edge start vertex strictly inside the rectangle"
  uint32_t inside = 0;
  bool vertex_in = false;
  const float f = B.front[a], r = B.rear[a], h = B.hw[a];
  for (int e = 0; e < B.n_road; ++e) {
    const float4 g = B.road_edge[e];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool cross = ((g.x > cy[k]) != (g.y > cy[k])) && (cx[k] < g.w * (cy[k] - g.x) + g.z);
      inside ^= (cross ? 1u : 0u) << k;
    }
    const float dx = g.z - t.x, dy = g.x - t.y;
    const float lx = dx * t.cs + dy * t.sn, ly = -dx * t.sn + dy * t.cs;
    vertex_in = (lx > -r && lx < f && fabsf(ly) < h) || vertex_in;
  }
  return inside != 0xFu || vertex_in;
}

__device__ __forceinline__ bool cw_at_goal(const CwBlob& B, const CwAgent& t, int a,
                                           const float (&cx)[4], const float (&cy)[4]) {
  if (!B.has_goal[a]) return false;
  const float* g = B.goal[a];
  bool in = true;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    in = in && cx[k] >= g[0] && cx[k] <= g[1] && cy[k] >= g[2] && cy[k] <= g[3];
  return in && t.th >= g[4] && t.th <= g[5];
}

// pose of agent j of my group, fetched with shuffles (executed by every lane of the warp)
struct CwPeer {
  float x, y, cs, sn, v, s;
  int lane;
  uint32_t flags;
};

__device__ __forceinline__ CwPeer cw_fetch(const CwAgent& t, int src) {
  CwPeer p;
  p.x = __shfl_sync(0xffffffffu, t.x, src);
  p.y = __shfl_sync(0xffffffffu, t.y, src);
  p.cs = __shfl_sync(0xffffffffu, t.cs, src);
  p.sn = __shfl_sync(0xffffffffu, t.sn, src);
  p.v = __shfl_sync(0xffffffffu, t.v, src);
  p.s = __shfl_sync(0xffffffffu, t.s, src);
  const int packed = __shfl_sync(0xffffffffu, (t.lane << 8) | static_cast<int>(t.flags), src);
  p.lane = packed >> 8;
  p.flags = static_cast<uint32_t>(packed) & 0xffu;
  return p;
}

// ---- passive agents: kCwPasWords words per slot in this thread's shared-memory column
constexpr int kCwPasWords = 13;

__device__ __forceinline__ void cw_pas_store(float* col, int slot, const CwAgent& t) {
  float* p = col + slot * kCwPasWords * kCwThreads;
  p[0 * kCwThreads] = t.x;  p[1 * kCwThreads] = t.y;  p[2 * kCwThreads] = t.th;  p[3 * kCwThreads] = t.v;
  p[4 * kCwThreads] = t.cs; p[5 * kCwThreads] = t.sn; p[6 * kCwThreads] = t.s;
  p[7 * kCwThreads] = t.ex; p[8 * kCwThreads] = t.ey; p[9 * kCwThreads] = t.eth; p[10 * kCwThreads] = t.ev;
  p[11 * kCwThreads] = __int_as_float((t.lane << 8) | static_cast<int>(t.flags));
  p[12 * kCwThreads] = t.lat;
}

__device__ __forceinline__ void cw_pas_load(const float* col, int slot, CwAgent& t) {
  const float* p = col + slot * kCwPasWords * kCwThreads;
  t.x = p[0 * kCwThreads];  t.y = p[1 * kCwThreads];  t.th = p[2 * kCwThreads];  t.v = p[3 * kCwThreads];
  t.cs = p[4 * kCwThreads]; t.sn = p[5 * kCwThreads]; t.s = p[6 * kCwThreads];
  t.ex = p[7 * kCwThreads]; t.ey = p[8 * kCwThreads]; t.eth = p[9 * kCwThreads]; t.ev = p[10 * kCwThreads];
  const int packed = __float_as_int(p[11 * kCwThreads]);
  t.lane = packed >> 8;
  t.flags = static_cast<uint32_t>(packed) & 0xffu;
  t.lat = p[12 * kCwThreads];
}

// pose of passive agent j of my rollout, read straight from its owner lane's shared-memory
// column (the lanes of a rollout sit next to each other, so the column is `owner - a` away);
// the owner's stores must be ordered before this by a __syncwarp()
__device__ __forceinline__ CwPeer cw_fetch_passive(const CwBlob& B, const float* pas, int j, int a, bool real) {
  const float* p = pas + (real ? static_cast<int>(B.pas_owner[j]) - a : 0) +
                   static_cast<int>(B.pas_slot[j]) * kCwPasWords * kCwThreads;
  CwPeer r;
  r.x = p[0 * kCwThreads];
  r.y = p[1 * kCwThreads];
  r.v = p[3 * kCwThreads];
  r.cs = p[4 * kCwThreads];
  r.sn = p[5 * kCwThreads];
  r.s = p[6 * kCwThreads];
  const int packed = __float_as_int(p[11 * kCwThreads]);
  r.lane = packed >> 8;
  r.flags = static_cast<uint32_t>(packed) & 0xffu;
  return r;
}

// separating-axis test of my rectangle against peer j's
__device__ __forceinline__ bool cw_rect_overlap(const CwBlob& B, const CwAgent& t, int a,
                                                const CwPeer& p, int j) {
  const float oi = (B.front[a] - B.rear[a]) * 0.5f, oj = (B.front[j] - B.rear[j]) * 0.5f;
  const float eix = (B.front[a] + B.rear[a]) * 0.5f, ejx = (B.front[j] + B.rear[j]) * 0.5f;
  const float eiy = B.hw[a], ejy = B.hw[j];
  const float dx = (p.x + oj * p.cs) - (t.x + oi * t.cs);
  const float dy = (p.y + oj * p.sn) - (t.y + oi * t.sn);
  // circumscribed circles apart: the rectangles cannot touch (exact early-out; most pairs)
  const float rr = B.rad[a] + B.rad[j];
  if (dx * dx + dy * dy > rr * rr) return false;
  const float cd = fabsf(t.cs * p.cs + t.sn * p.sn), sd = fabsf(t.sn * p.cs - t.cs * p.sn);
  bool hit = fabsf(dx * t.cs + dy * t.sn) <= eix + ejx * cd + ejy * sd;
  hit = (fabsf(-dx * t.sn + dy * t.cs) <= eiy + ejx * sd + ejy * cd) && hit;
  hit = (fabsf(dx * p.cs + dy * p.sn) <= ejx + eix * cd + eiy * sd) && hit;
  hit = (fabsf(-dx * p.sn + dy * p.cs) <= ejy + eix * sd + eiy * cd) && hit;
  return hit;
}

// Everything of Execute that follows the kinematic step, for the lane's own agent:
// collision / drivable area / goal, labels, rule automata, reward vector, validity and
// the state's terminal flag.  Must be called by all 32 lanes (shuffles inside).
//   a, base : my agent index and the first lane of my group; real: lane maps to an agent
//   live    : my group executes this step
//   in_*    : my agent's v, theta, flags BEFORE the kinematic step
//   q, viol : this thread's column of the per-slot automaton state / violation counters
//   rules_only : MvmctsState::EvaluateRules() (:186-197) on the state as it is -- labels and
//                automaton transitions of every MCTS agent, no other reward term, no invalidation
template <bool WITH_LABELS, bool PASSIVE>
__device__ __forceinline__ void cw_evaluate(const CwBlob& B, CwAgent& t, int a, int base, bool real,
                                            bool live, float in_v, float in_th, uint32_t in_flags,
                                            int horizon_after, uint8_t* q, uint8_t* viol,
                                            int col_stride, const float* pas, CwStepOut& o,
                                            bool rules_only = false) {
  const int A = B.A, M = B.M, V = B.V;
  const bool present = real && (t.flags & BRM_CW_PRESENT);
  const bool evaluated = live && present && (in_flags & BRM_CW_VALID) && a < M;
  CwVec r = {0.f, 0.f, 0.f, 0.f};
  o.w[0] = o.w[1] = o.w[2] = o.w[3] = 0u;

  bool oom = false, goal = false;
  if (evaluated || (live && present && a == 0)) {
    float cx[4], cy[4];
    cw_corners(B, t, a, cx, cy);
    oom = cw_out_of_map(B, t, a, cx, cy);
    goal = cw_at_goal(B, t, a, cx, cy);
  }
  // pairwise predicates against the agents of my rollout
  bool coll = false;
  int jf = -1;
  float gf = 0.f, vf = 0.f;
  uint32_t merged_mask = 0, near_mask = 0;
  const float my_len = (t.lane >= 0) ? B.lanes[t.lane].length : 0.f;
  const int my_succ = (t.lane >= 0) ? B.lanes[t.lane].succ : -2;
  auto pair = [&](int j, const CwPeer& p) {
    if (j == a || !(p.flags & BRM_CW_PRESENT) || !present) return;
    coll = cw_rect_overlap(B, t, a, p, j) || coll;
    // gap along my corridor: own lane, then its successor
    if (t.lane >= 0 && p.lane >= 0) {
      float g = 0.f;
      bool has = false;
      if (p.lane == t.lane) {
        g = p.s - t.s;
        has = true;
      } else if (my_succ == p.lane) {
        g = (my_len - t.s) + p.s;
        has = true;
      }
      if (has && g > 0.f && (jf < 0 || g < gf)) {
        jf = j;
        gf = g;
        vf = p.v;
      }
    }
    if (p.lane >= 0 && p.s >= B.lanes[p.lane].s_point) merged_mask |= 1u << j;
    const float dx = p.x - t.x, dy = p.y - t.y;
    if (dx * dx + dy * dy < B.near2) near_mask |= 1u << j;
  };
  // MCTS agents sit in the registers of the lanes of my group
  for (int j = 0; j < M; ++j) pair(j, cw_fetch(t, base + j));
  if (PASSIVE) {
    // the other agents in the shared-memory slots of those lanes (ascending j: ties in the
    // preceding-agent search resolve as in one loop over all agents)
    __syncwarp();
    for (int j = M; j < A; ++j) pair(j, cw_fetch_passive(B, pas, j, a, real));
    __syncwarp();
  }

  if (evaluated) {
    if (!rules_only) {
      r.c0 += (coll ? 1.f : 0.f) * B.w_coll;  // :91-92
      r.c0 += (oom ? 1.f : 0.f) * B.w_oom;    // :94-96
      // GetActionCost (:202-248)
      const float acc = (t.v - in_v) * B.inv_dt;
      float cost = B.w_acc * acc * acc * B.dt;
      const float theta_dot = (t.th - in_th) * B.inv_dt;
      const float avg_vel = 0.5f * acc * B.dt + in_v;
      const float a_lat = theta_dot * avg_vel;
      cost += B.w_rad * a_lat * a_lat * B.dt;
      cost += B.w_vel * fabsf(in_v - B.v_des) * B.dt;
      if (t.lane >= 0) cost += B.w_lane * fabsf(t.lat) * B.dt;
      // PotentialReward (:250-264)
      const float pot_new = -B.w_pot * B.dt * fabsf(t.v - B.v_des);
      const float pot_cur = -B.w_pot * B.dt * fabsf(in_v - B.v_des);
      const float pot = B.gamma * pot_new - pot_cur;
      vec_add_at(r, V - 1, cost);
      vec_add_at(r, V - 1, pot);
    }
    // labels (EvaluateRules, :165-185)
    uint32_t w0 = 0;
    if (coll) w0 |= 1u << BRM_CW_LABEL_COLLISION_EGO;
    if (oom) w0 |= 1u << BRM_CW_LABEL_COLLISION_CORRIDOR;
    if (goal) w0 |= 1u << BRM_CW_LABEL_GOAL_REACHED;
    bool safe = true;
    if (jf >= 0) {
      const float dist = gf - (B.front[a] + B.rear[jf]);
      const float dsafe = t.v * B.sd_delta + (t.v * t.v) * B.sd_ar - (vf * vf) * B.sd_af;
      safe = dist >= dsafe;
    }
    if (safe) w0 |= 1u << BRM_CW_LABEL_SAFE_DISTANCE_FRONT;
    if (t.lane >= 0) {
      const CwLaneHdr L = B.lanes[t.lane];
      if (t.s >= L.s_point) w0 |= 1u << BRM_CW_LABEL_MERGED_E;
      if (L.flags & 1) w0 |= 1u << BRM_CW_LABEL_RIGHTMOST_LANE;
      w0 |= 1u << BRM_CW_LABEL_ON_ROAD;
    }
    const uint32_t front_mask = jf >= 0 ? (1u << jf) : 0u;
    if (WITH_LABELS) {
      o.w[0] = w0; o.w[1] = merged_mask; o.w[2] = front_mask; o.w[3] = near_mask;
    }
    if (rules_only || !(B.ego_only && a != 0)) {  // :104-108
      CwVec rr = {0.f, 0.f, 0.f, 0.f};
      for (int sl = 0; sl < B.R; ++sl) {
        // one 16-byte record per slot: x = label-word bit read by each of the 4 APs (a byte each),
        // y = n_aps | priority << 8 | init << 16 | on_violation << 24,
        // z = rule | (rank + 1) << 8, w = weight
        const uint4 rec = B.slot_rec[sl];
        const int rank = static_cast<int>((rec.z >> 8) & 0xffu) - 1;
        const int other = rank < 0 ? 0 : (rank < a ? rank : rank + 1);
        const uint32_t n_aps = rec.y & 0xffu;
        const uint32_t cur = q[sl * col_stride];
        // absorbing state without violations (e.g. the ZIP monitor once its antecedent has
        // failed): nothing can change any more, skip the letter and the table
        if ((rec.z >> (16 + cur)) & 1u) continue;
        // label word of this slot: ego kinds in bits 0..7, the bound agent's kinds in 8..10
        const uint32_t W = w0 | (((merged_mask >> other) & 1u) << BRM_CW_LABEL_MERGED_X) |
                           (((front_mask >> other) & 1u) << BRM_CW_LABEL_IN_DIRECT_FRONT_X) |
                           (((near_mask >> other) & 1u) << BRM_CW_LABEL_NEAR_X);
        // letter bit k = AP k (the blob's tables are stored in this order; unused APs read the
        // always-zero bit 31)
        uint32_t letter = 0;
#pragma unroll
        for (uint32_t k = 0; k < BRM_MAX_APS; ++k)
          letter |= ((W >> ((rec.x >> (8 * k)) & 31u)) & 1u) << k;
        uint32_t nxt = B.rules[rec.z & 0xffu].trans[(cur << n_aps) + letter];
        if (nxt == BRM_VIOLATION) {
          viol[sl * col_stride] += 1;
          nxt = (rec.y >> 24) ? cur : ((rec.y >> 16) & 0xffu);
          const uint32_t pr = (rec.y >> 8) & 0xffu;
          vec_add_at(rr, static_cast<int>(pr), __uint_as_float(rec.w));
        }
        if (nxt != cur) q[sl * col_stride] = static_cast<uint8_t>(nxt);
      }
      r.c0 += rr.c0; r.c1 += rr.c1; r.c2 += rr.c2; r.c3 += rr.c3;
    }
    if (goal && !rules_only) vec_add_at(r, V - 1, B.goal_reward);  // :110-113
    if (coll && !rules_only) t.flags &= ~static_cast<uint32_t>(BRM_CW_VALID);  // :116-120
  }
  // CheckTerminal of the new state (:266-285), decided by the ego lane of the group
  const bool ego_term = !present || horizon_after == 0 || oom || coll || goal;
  o.terminal = __shfl_sync(0xffffffffu, ego_term ? 1 : 0, base) != 0;
  o.r[0] = r.c0; o.r[1] = r.c1; o.r[2] = r.c2; o.r[3] = r.c3;
}

// GetTerminalReward (:148-164) for my agent
__device__ __forceinline__ void cw_terminal_reward(const CwBlob& B, const CwAgent& t, int a,
                                                   bool real, const uint8_t* q, int col_stride,
                                                   float (&r)[BRM_MAX_REWARD]) {
#pragma unroll
  for (int v = 0; v < BRM_MAX_REWARD; ++v) r[v] = 0.f;
  if (!real || a >= B.M || !(t.flags & BRM_CW_PRESENT)) return;
  CwVec acc = {0.f, 0.f, 0.f, 0.f};
  for (int sl = 0; sl < B.R; ++sl) {
    const brm_rule& rule = B.rules[B.slot_rule[sl]];
    const float pen = rule.accepting[q[sl * col_stride]] ? 0.f : rule.weight;
    vec_add_at(acc, rule.priority, pen);
  }
  r[0] = acc.c0; r[1] = acc.c1; r[2] = acc.c2; r[3] = acc.c3;
}

}  // namespace brm
