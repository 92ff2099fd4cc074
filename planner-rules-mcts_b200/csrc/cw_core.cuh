// Continuous-world step arithmetic: MvmctsState::Execute
// (/root/reference/src/rules_mcts_state.cpp:57-128) cut into per-agent work items that
// operate on a pool of rollouts held in shared memory.
//
//   phase 1, item = (rollout, agent that moves):  draw the action (Philox), single-track
//            kinematics, drivable-area / goal test, lane projection, action cost + potential
//   phase 2, item = (rollout, MCTS agent):  collision and label predicates against the other
//            agents of the rollout, rule automata, reward vector, validity, terminal flag
//   init / finalize items load a leaf into a pool slot and fold a finished rollout into the
//            per-leaf sums.
//
// All state arithmetic is FP64, the reference's type: labels, automaton states and violation
// counts have to equal the reference's double arithmetic bit for bit, and an fp32 state
// cannot guarantee that within 1e-5 m of a threshold.  fp32 is only used where it cannot
// change a result: conservative culling boxes of the lane projection.
//
// The code is host/device portable (BRM_HD) and uses explicit fma() -- the library is
// compiled with --fmad=false -- so that tests/cpp/cw_emul.cpp can run exactly this
// arithmetic on the CPU (bit-identical to the GPU) against the oracle without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "brm.h"

#if defined(__CUDACC__)
#define BRM_HD __host__ __device__ __forceinline__
#define BRM_HD_NOINLINE inline __host__ __device__ __noinline__  // one copy in the kernel: the code has to fit the instruction cache
#else
#define BRM_HD inline
#define BRM_HD_NOINLINE inline
#endif

// loops whose unrolling costs more instruction-cache misses than it gains
#if defined(__CUDA_ARCH__) && !defined(BRM_CW_BIG_FRENET)
#define BRM_LOOP_SMALL _Pragma("unroll 1")
#else
#define BRM_LOOP_SMALL
#endif
#if defined(__CUDA_ARCH__)
#define BRM_LOOP_EULER _Pragma("unroll 2")
#else
#define BRM_LOOP_EULER
#endif

namespace brm {

// ------------------------------------------------------------------ scenario blob
// lane segments are stored field by field (x0[], y0[], dx[], dy[], inv_len2[], len[], s0[], each
// n_seg_total long): the lanes of a warp project different points onto different segments, and
// 8-byte strides spread those reads over the shared-memory banks (64-byte records would put
// every segment on the same two)
enum { SEG_X0 = 0, SEG_Y0, SEG_DX, SEG_DY, SEG_INV, SEG_LEN, SEG_S0, SEG_FIELDS };

constexpr int kCwBuckets = 64;  // buckets of a lane's segment index (cw_frenet)

struct alignas(16) CwLaneHdr {
  int32_t seg_off, n_seg, succ, flags;
  double hw2, length, s_point, _pad;
  // fp32, conservative: bounding box of the centre line grown by the half width; the lane's
  // extent along its longer axis cut into kCwBuckets buckets (origin, 1 / width)
  float gx0, gy0, gx1, gy1;
  float b_org, b_inv;
  int32_t b_axis, _pi;
};

struct alignas(16) CwEdge {
  double ya, yb, xa, m;  // m = (xb - xa) / (yb - ya)
};

struct alignas(16) CwBox {
  float x0, y0, x1, y1;
};

struct alignas(16) CwSlotRec {
  uint32_t x;  // label-word bit read by each of the 4 APs (a byte each; 31 = always zero)
  uint32_t y;  // n_aps | priority << 8 | init << 16 | on_violation << 24
  uint32_t z;  // rule | (rank + 1) << 8 | absorbing-state mask << 16
  float w;     // weight
};

constexpr int kCwMaxSegs = BRM_CW_MAX_LANES * (BRM_CW_MAX_PTS - 1);
constexpr int kCwGrid = 64;  // cells per axis of the drivable-area grid
// drivable-area grid: 4 bits per cell
enum : uint32_t {
  CW_CELL_INSIDE = 1,   // the whole cell (and a margin) lies inside the road polygon
  CW_CELL_OUTSIDE = 2,  // the whole cell (and a margin) lies outside
  CW_CELL_VERTEX = 4,   // a road vertex lies within the largest agent's reach of the cell
};

struct alignas(16) CwBlob {
  int32_t A, M, V, R, n_lanes, n_road, n_rules, n_sub;
  int32_t ego_only, total_bytes, _pb, remove_oom;
  int32_t n_seg_total, _pi0, _pi1, _pi2;
  double w_coll, w_oom, w_pot, w_acc, w_rad, w_vel, w_lane, dt;
  double v_des, gamma, goal_reward, h_sub, h_last, sd_delta, sd_ar, sd_af;
  double near2, c2, t_total, inv_dt;  // c2, t_total: closed form of the Euler sub-steps without steering
  double front[BRM_CW_MAX_AGENTS], rear[BRM_CW_MAX_AGENTS], hw[BRM_CW_MAX_AGENTS];
  double off[BRM_CW_MAX_AGENTS], ext[BRM_CW_MAX_AGENTS];  // (front - rear) / 2, (front + rear) / 2
  double rad[BRM_CW_MAX_AGENTS];  // circumscribed radius of the rectangle about its centre (a hair more)
  double goal[BRM_CW_MAX_AGENTS][6];
  double prim_a[BRM_CW_MAX_AGENTS][BRM_CW_MAX_PRIMS];
  double prim_k[BRM_CW_MAX_AGENTS][BRM_CW_MAX_PRIMS];  // tan(delta) / wheel_base
  uint8_t has_goal[BRM_CW_MAX_AGENTS], n_prims[BRM_CW_MAX_AGENTS];
  CwSlotRec slot_rec[BRM_CW_MAX_SLOTS];
  uint8_t slot_rule[BRM_CW_MAX_SLOTS];
  brm_rule rules[BRM_MAX_RULES];  // trans stored with AP k of the letter at bit k
  CwEdge road_edge[BRM_CW_MAX_ROAD];
  // nearest-segment search (cw_frenet): per lane and bucket of its longer axis the range
  // [lo, hi] (lo | hi << 8; lo > hi: none) of the segments that come within the half width of
  // any point of the bucket -- built with a slack that covers the fp32 evaluation of the bucket
  // index for coordinates up to BRM_CW_MAX_COORD
  uint16_t bucket[BRM_CW_MAX_LANES][kCwBuckets];
  // drivable-area test (cw_classify_shape): a grid over the road polygon's bounding box, 4
  // bits per cell (CW_CELL_*), 8 cells per word; origin and 1 / cell size
  float grid_x0, grid_y0, grid_ix, grid_iy;
  uint32_t grid[kCwGrid * kCwGrid / 8];
  CwLaneHdr lanes[BRM_CW_MAX_LANES];
  double segs[SEG_FIELDS * kCwMaxSegs];  // last: only the used part (SEG_FIELDS * n_seg_total) is copied
};
static_assert(sizeof(CwBlob) % 16 == 0, "blob must be a multiple of 16 bytes");
static_assert(offsetof(CwBlob, segs) % 16 == 0, "segment array must be 16-byte aligned");

// ------------------------------------------------------------------ the pool
// P rollout slots of A agents (M of them MCTS agents) in structure-of-arrays form; in the
// kernels these arrays live in shared memory.
enum : uint8_t {
  CW_BIT_OOM = 1,       // drivable-area test of this step: shape not within the road polygon
  CW_BIT_GOAL = 2,      // shape within the goal
  CW_BIT_VANISHED = 4,  // left the map during this step (remove_agents_out_of_map)
  CW_BIT_PENDING = 8,   // phase 1 could not tell the drivable-area test from the grid: cw_exact_shape_item
  CW_BIT_BEYOND = 16,   // on a lane, at or beyond its merge point (s >= s_point): merged_e / merged_x#j
  CW_BIT_PENDING_V = 32,  // with PENDING: the corners are certainly inside, only the road vertices are left to test
};

// Records are padded to an odd number of 8-byte words, so that the lanes of a warp, which
// work on neighbouring slots, read a field from different banks.
constexpr int kAgWords = 9;  // agent record: x, y, th, v, cs, sn, s, lat, cost of the action just taken
struct CwMeta {              // 4 bytes per agent, in an array of their own (p.meta): consecutive
  int16_t lane;              // slots sit in consecutive banks (inside the 9-word records every
  uint8_t flags;             // 32-bit access would meet the lane 16 slots away in the same bank)
  uint8_t bits;              // lane of the current position (-1: none); BRM_CW_PRESENT | VALID; CW_BIT_*
};
static_assert(sizeof(CwMeta) == 4, "CwMeta is one 32-bit word");
enum { AG_X = 0, AG_Y, AG_TH, AG_V, AG_CS, AG_SN, AG_S, AG_LAT, AG_COST };
// MCTS agent record, in 8-byte words: acc[V], rng[4 x u32], then u32 live, u32 qw[QW], u8 viol[R]
enum { MC_ACC = 0 };
// rollout header, 6 x int32: flat, leaf, k, horizon, nsteps, state bytes (busy, done, live, rflag)
enum { H_FLAT = 0, H_LEAF, H_K, H_HORIZON, H_NSTEPS, H_STATE, H_WORDS = 6 };
enum { HS_BUSY = 0, HS_DONE = 1, HS_LIVE = 2, HS_RFLAG = 3 };

struct CwPool {
  int32_t P, A, M, R, V, QW, mw;  // QW = words of packed automaton states per agent; mw = words per MCTS record
  double* ag;    // [P * A][kAgWords]
  uint32_t* meta;  // [P * A] CwMeta
  double* mc;    // [P * M][mw]
  double* disc;  // [P] discount of the current step
  int32_t* hdr;  // [H_WORDS][P]
};

BRM_HD int cw_mc_words(int R, int V) {
  const int QW = (R + 7) / 8;
  int w = V + 2 + (4 * (1 + QW) + R + 7) / 8;
  return (w & 1) ? w : w + 1;
}

BRM_HD size_t cw_pool_bytes(int P, int A, int M, int R, int V) {
  size_t b = static_cast<size_t>(P) * A * kAgWords * sizeof(double);
  b += static_cast<size_t>(P) * M * cw_mc_words(R, V) * sizeof(double);
  b += static_cast<size_t>(P) * sizeof(double);
  b += (static_cast<size_t>(P) * A * sizeof(uint32_t) + 7) & ~static_cast<size_t>(7);
  b += static_cast<size_t>(P) * H_WORDS * sizeof(int32_t);
  return (b + 15) & ~static_cast<size_t>(15);
}

// carve the arrays out of `mem` (16-byte aligned, cw_pool_bytes large)
BRM_HD void cw_pool_carve(CwPool& p, void* mem, int P, int A, int M, int R, int V) {
  p.P = P; p.A = A; p.M = M; p.R = R; p.V = V; p.QW = (R + 7) / 8;
  p.mw = cw_mc_words(R, V);
  double* d = static_cast<double*>(mem);
  p.ag = d; d += static_cast<size_t>(P) * A * kAgWords;
  p.mc = d; d += static_cast<size_t>(P) * M * p.mw;
  p.disc = d; d += P;
  p.meta = reinterpret_cast<uint32_t*>(d); d += (static_cast<size_t>(P) * A + 1) / 2;
  p.hdr = reinterpret_cast<int32_t*>(d);
}

// Slot of agent j of rollout r.  Agent-major: the work lists group items by agent, so the lanes of
// a warp hold the same agent of consecutive rollouts -- stored next to each other (one odd record
// stride apart), they read a field from different banks; rollout-major storage would put them
// A records apart and, for A = 4, eight lanes on every bank.
BRM_HD int cw_slot(const CwPool& p, int r, int j) { return j * p.P + r; }
BRM_HD int cw_mslot(const CwPool& p, int r, int a) { return a * p.P + r; }
BRM_HD double* cw_ag(const CwPool& p, int s) { return p.ag + s * kAgWords; }
BRM_HD CwMeta* cw_meta(const CwPool& p, int s) { return reinterpret_cast<CwMeta*>(p.meta + s); }
BRM_HD double* cw_mc(const CwPool& p, int sm) { return p.mc + sm * p.mw; }
BRM_HD uint32_t* cw_mc_rng(const CwPool& p, int sm) { return reinterpret_cast<uint32_t*>(cw_mc(p, sm) + p.V); }
BRM_HD uint32_t* cw_mc_live(const CwPool& p, int sm) { return reinterpret_cast<uint32_t*>(cw_mc(p, sm) + p.V + 2); }
BRM_HD uint32_t* cw_mc_qw(const CwPool& p, int sm) { return cw_mc_live(p, sm) + 1; }
BRM_HD uint8_t* cw_mc_viol(const CwPool& p, int sm) { return reinterpret_cast<uint8_t*>(cw_mc_qw(p, sm) + p.QW); }
// rollout headers are stored field by field ([H_WORDS][P]): S1 maps a lane to a rollout slot, and
// a record per rollout (8 words) would put every fourth lane on the same bank
struct CwHdrRef {
  int32_t* base;  // field 0 of this rollout
  int32_t P;
  BRM_HD int32_t& operator[](int field) const { return base[field * P]; }
};
BRM_HD CwHdrRef cw_hdr(const CwPool& p, int r) { return CwHdrRef{p.hdr + r, p.P}; }
BRM_HD uint8_t* cw_hstate(const CwPool& p, int r) { return reinterpret_cast<uint8_t*>(p.hdr + H_STATE * p.P + r); }

// ------------------------------------------------------------------ small math
BRM_HD double cw_fma(double a, double b, double c) { return fma(a, b, c); }

// sin and cos of x for moderate |x| (headings): Cody-Waite reduction by pi/2 in three parts,
// then the fdlibm kernel polynomials on [-pi/4, pi/4].  Below 2 ulp; identical on host and
// device (only fma, +, *).
BRM_HD void cw_sincos(double x, double* sn, double* cs) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: (t + magic) - magic = rint(t)
  const double n = (x * 0.63661977236758134308 + kMagic) - kMagic;
  double r = cw_fma(-n, 1.5707963267948966, x);
  r = cw_fma(-n, 6.123233995736766e-17, r);
  r = cw_fma(-n, -1.4973849048591698e-33, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = cw_fma(ps, z, -2.50507602534068634195e-08);
  ps = cw_fma(ps, z, 2.75573137070700676789e-06);
  ps = cw_fma(ps, z, -1.98412698298579493134e-04);
  ps = cw_fma(ps, z, 8.33333333332248946124e-03);
  ps = cw_fma(ps, z, -1.66666666666666324348e-01);
  ps = cw_fma(ps * z, r, r);
  double pc = -1.13596475577881948265e-11;
  pc = cw_fma(pc, z, 2.08757232129817482790e-09);
  pc = cw_fma(pc, z, -2.75573143513906633035e-07);
  pc = cw_fma(pc, z, 2.48015872894767294178e-05);
  pc = cw_fma(pc, z, -1.38888888888741095749e-03);
  pc = cw_fma(pc, z, 4.16666666666666019037e-02);
  pc = cw_fma(pc * z, z, cw_fma(-0.5, z, 1.0));
  const int q = static_cast<int>(static_cast<long long>(n)) & 3;
  const double s0 = (q & 1) ? pc : ps, c0 = (q & 1) ? ps : pc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}

// sin and cos of a small angle (|d| <= 1/8): Taylor polynomials, error below 1e-20
BRM_HD void cw_sincos_small(double d, double* sn, double* cs) {
  const double z = d * d;
  double ps = -1.0 / 39916800.0;
  ps = cw_fma(ps, z, 1.0 / 362880.0);
  ps = cw_fma(ps, z, -1.0 / 5040.0);
  ps = cw_fma(ps, z, 1.0 / 120.0);
  ps = cw_fma(ps, z, -1.0 / 6.0);
  *sn = cw_fma(ps * z, d, d);
  double pc = 1.0 / 479001600.0;
  pc = cw_fma(pc, z, -1.0 / 3628800.0);
  pc = cw_fma(pc, z, 1.0 / 40320.0);
  pc = cw_fma(pc, z, -1.0 / 720.0);
  pc = cw_fma(pc, z, 1.0 / 24.0);
  pc = cw_fma(pc, z, -0.5);
  *cs = cw_fma(pc, z, 1.0);
}

// ------------------------------------------------------------------ Philox4x32-10
BRM_HD uint32_t cw_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

// counter = (rollout_lo, rollout_hi, step >> 2, agent), key = seed; word step & 3 serves `step`
BRM_HD void cw_philox_block(uint64_t seed, uint64_t rollout, uint32_t step, uint32_t agent, uint32_t* out) {
  uint32_t c0 = static_cast<uint32_t>(rollout), c1 = static_cast<uint32_t>(rollout >> 32);
  uint32_t c2 = step >> 2, c3 = agent;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = cw_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = cw_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ------------------------------------------------------------------ kinematics
struct CwKin {
  double x, y, th, v, cs, sn;
};

// bark SingleTrackModel under BehaviorMPContinuousActions (un-vendored): n_sub explicit Euler
// sub-steps  x += h (v cos th), y += h (v sin th), th += h (v k), v += h a.
// The speed and the heading never depend on the position, so their recurrences are carried out
// operation by operation like the reference's loop (no fused multiply-add): v and th are
// bit-identical to the double-precision loop, always.  Without steering the heading is
// constant and the position follows the same recurrence literally -- round-number scenarios
// (straight roads, speeds like 5.0) put agents EXACTLY on thresholds (8.0 m apart), and only
// the same sequence of roundings decides those the same way.  With steering the heading of
// sub-step n+1 is obtained by rotating through the small angle h v_n k instead of a fresh
// sine / cosine per sub-step (positions agree to ~1e-15; they are generic numbers there).
BRM_HD void cw_integrate(const CwBlob& B, CwKin& t, double acc, double kap) {
  const int n_sub = B.n_sub;
  double x = t.x, y = t.y, v = t.v;
  if (kap == 0.0) {
    // n_sub - 1 sub-steps of h_sub and a last one of h_last (the loop is peeled: a select of h
    // per sub-step cost as much as the arithmetic); h_sub * acc is the same product every time
    const double cs = t.cs, sn = t.sn, hs = B.h_sub, hl = B.h_last, dv = hs * acc;
    BRM_LOOP_EULER
    for (int n = 1; n < n_sub; ++n) {
      x = x + hs * (v * cs);
      y = y + hs * (v * sn);
      v = v + dv;
    }
    x = x + hl * (v * cs);
    y = y + hl * (v * sn);
    v = v + hl * acc;
  } else {
    double dx = 0.0, dy = 0.0, th = t.th, c = t.cs, s = t.sn;
    BRM_LOOP_SMALL
    for (int n = 0; n < n_sub; ++n) {
      const double h = (n == n_sub - 1) ? B.h_last : B.h_sub;
      const double hv = h * v;
      dx = cw_fma(hv, c, dx);
      dy = cw_fma(hv, s, dy);
      // heading of the next sub-step: th += h (v k), the pose rotated by the same angle
      const double dth = h * (v * kap);
      th = th + dth;
      double sd, cd;
      if (fabs(dth) <= 0.125)
        cw_sincos_small(dth, &sd, &cd);
      else
        cw_sincos(dth, &sd, &cd);
      const double c1 = cw_fma(c, cd, -(s * sd));
      s = cw_fma(s, cd, c * sd);
      c = c1;
      v = v + h * acc;
    }
    x += dx;
    y += dy;
    t.th = th;
    cw_sincos(th, &t.sn, &t.cs);
  }
  t.x = x;
  t.y = y;
  t.v = v;
}

// ------------------------------------------------------------------ lane projection
struct CwFrenet {
  int lane;
  double s, lat;
};

// Nearest point on every lane centre line; the first lane that contains the point wins
// (the brute-force search over all segments of the oracle).  What is skipped cannot change
// the answer: a lane whose grown bounding box does not contain the point, and on a lane every
// segment outside the bucket range of the point's coordinate -- a segment within the half
// width of the point is in that range, and if none is, the lane does not contain the point.
// Distances themselves are FP64.
BRM_HD_NOINLINE CwFrenet cw_frenet(const CwBlob& B, double x, double y) {
  CwFrenet f;
  f.lane = -1;
  f.s = 0.0;
  f.lat = 0.0;
  const float xf = static_cast<float>(x), yf = static_cast<float>(y);
  const int ns = B.n_seg_total;
  const double* X0 = B.segs + SEG_X0 * ns;
  const double* Y0 = B.segs + SEG_Y0 * ns;
  const double* DX = B.segs + SEG_DX * ns;
  const double* DY = B.segs + SEG_DY * ns;
  const double* INV = B.segs + SEG_INV * ns;
  // candidate lanes first (uniform loop), then one candidate per round: the lanes of a warp
  // look at different road lanes, but they all do it at the same time
  uint32_t cand = 0u;
  const int n_lanes = B.n_lanes;
  BRM_LOOP_SMALL
  for (int l = 0; l < n_lanes; ++l) {
    const CwLaneHdr& L = B.lanes[l];
    if (xf >= L.gx0 && xf <= L.gx1 && yf >= L.gy0 && yf <= L.gy1) cand |= 1u << l;
  }
  while (cand) {
#if defined(__CUDA_ARCH__)
    const int l = __ffs(static_cast<int>(cand)) - 1;
#else
    const int l = __builtin_ctz(cand);
#endif
    cand &= cand - 1u;
    const CwLaneHdr& L = B.lanes[l];
    const float fb = ((L.b_axis ? yf : xf) - L.b_org) * L.b_inv;
    int b = static_cast<int>(fb);
    b = b < 0 ? 0 : (b > kCwBuckets - 1 ? kCwBuckets - 1 : b);
    const uint32_t range = B.bucket[l][b];
    const int k_lo = static_cast<int>(range & 0xffu), k_hi = static_cast<int>(range >> 8);
    const int seg_off = L.seg_off;
    int bk = 0;
    double best = 1e300, btu = 0.0;
    BRM_LOOP_SMALL
    for (int k = k_lo; k <= k_hi; ++k) {
      const int g = seg_off + k;
      const double wx = x - X0[g], wy = y - Y0[g];
      const double sdx = DX[g], sdy = DY[g];
      const double tu = cw_fma(wx, sdx, wy * sdy) * INV[g];
      const double tc = fmin(fmax(tu, 0.0), 1.0);
      const double ex = cw_fma(-tc, sdx, wx), ey = cw_fma(-tc, sdy, wy);
      const double d2 = cw_fma(ex, ex, ey * ey);
      if (d2 < best) {
        best = d2;
        bk = k;
        btu = tu;
      }
    }
    if (best <= L.hw2) {
      const bool interior = !(bk == 0 && btu < 0.0) && !(bk == L.n_seg - 1 && btu > 1.0);
      if (interior) {
        const int g = seg_off + bk;
        const double bt = fmin(fmax(btu, 0.0), 1.0);
        const double wx = x - X0[g], wy = y - Y0[g];
        const double cross = cw_fma(DX[g], wy, -(DY[g] * wx));
        f.lane = l;
        f.s = cw_fma(bt, B.segs[SEG_LEN * ns + g], B.segs[SEG_S0 * ns + g]);
        const double dd = sqrt(best);
        f.lat = cross < 0.0 ? -dd : dd;
        return f;
      }
    }
  }
  return f;
}

// the beyond-point label of a projected agent (EgoBeyondPoint / AgentBeyondPoint label functions)
BRM_HD uint8_t cw_beyond_bit(const CwBlob& B, const CwFrenet& f) {
  return (f.lane >= 0 && f.s >= B.lanes[f.lane].s_point) ? CW_BIT_BEYOND : 0;
}

// ------------------------------------------------------------------ shape predicates
BRM_HD void cw_corners(const CwBlob& B, const CwKin& t, int a, double (&cx)[4], double (&cy)[4]) {
  const double f = B.front[a], r = B.rear[a], h = B.hw[a];
  const double lx[4] = {f, f, -r, -r};
  const double ly[4] = {h, -h, -h, h};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    cx[k] = t.x + cw_fma(lx[k], t.cs, -(ly[k] * t.sn));
    cy[k] = t.y + cw_fma(lx[k], t.sn, ly[k] * t.cs);
  }
}

// EvaluatorDrivableArea restated: a corner outside the road polygon (crossing number), or a
// road vertex strictly inside the rectangle
BRM_HD bool cw_out_of_map_corners(const CwBlob& B, const CwKin& t, int a, const double (&cx)[4], const double (&cy)[4]) {
  uint32_t inside = 0;
  bool vertex_in = false;
  const double f = B.front[a], r = B.rear[a], h = B.hw[a];
  const int n_road = B.n_road;
  for (int e = 0; e < n_road; ++e) {
    const CwEdge g = B.road_edge[e];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
      const bool cross = ((g.ya > cy[k]) != (g.yb > cy[k])) && (cx[k] < cw_fma(g.m, cy[k] - g.ya, g.xa));
      inside ^= (cross ? 1u : 0u) << k;
    }
    const double dx = g.xa - t.x, dy = g.ya - t.y;
    const double lx = cw_fma(dx, t.cs, dy * t.sn), ly = cw_fma(dy, t.cs, -(dx * t.sn));
    vertex_in = (lx > -r && lx < f && fabs(ly) < h) || vertex_in;
  }
  return inside != 0xFu || vertex_in;
}

// the second half of the exact test alone (a road vertex strictly inside the rectangle), for
// shapes whose corners the grid has already placed inside the polygon; same arithmetic
BRM_HD bool cw_vertex_in_shape(const CwBlob& B, const CwKin& t, int a) {
  bool vertex_in = false;
  const double f = B.front[a], r = B.rear[a], h = B.hw[a];
  const int n_road = B.n_road;
  BRM_LOOP_SMALL
  for (int e = 0; e < n_road; ++e) {
    const double dx = B.road_edge[e].xa - t.x, dy = B.road_edge[e].ya - t.y;
    const double lx = cw_fma(dx, t.cs, dy * t.sn), ly = cw_fma(dy, t.cs, -(dx * t.sn));
    vertex_in = (lx > -r && lx < f && fabs(ly) < h) || vertex_in;
  }
  return vertex_in;
}

// the exact test from a pose (everything by value: one out-of-line copy in the kernels)
BRM_HD_NOINLINE bool cw_out_of_map_pose(const CwBlob& B, double x, double y, double cs, double sn, int a) {
  CwKin t;
  t.x = x; t.y = y; t.th = 0.0; t.v = 0.0; t.cs = cs; t.sn = sn;
  double cx[4], cy[4];
  cw_corners(B, t, a, cx, cy);
  return cw_out_of_map_corners(B, t, a, cx, cy);
}
BRM_HD bool cw_out_of_map(const CwBlob& B, const CwKin& t, int a, const double (&)[4], const double (&)[4]) {
  return cw_out_of_map_pose(B, t.x, t.y, t.cs, t.sn, a);
}

// Drivable-area test through the grid: 1 = certainly out of the map, 0 = certainly inside,
// -1 = a corner too close to the polygon's boundary to tell (the caller runs cw_out_of_map),
// -2 = corners certainly inside but a road vertex within reach (the caller runs cw_vertex_in_shape).  A corner in an OUTSIDE cell is outside the polygon; four corners in INSIDE
// cells with no road vertex within reach of the centre's cell leave nothing that could make
// the exact test true.
BRM_HD int cw_classify_shape(const CwBlob& B, const CwKin& t, const double (&cx)[4], const double (&cy)[4]) {
  auto cell = [&](double px, double py) -> uint32_t {
    const float fx = (static_cast<float>(px) - B.grid_x0) * B.grid_ix;
    const float fy = (static_cast<float>(py) - B.grid_y0) * B.grid_iy;
    if (!(fx >= 0.f && fx < static_cast<float>(kCwGrid) && fy >= 0.f && fy < static_cast<float>(kCwGrid)))
      return CW_CELL_OUTSIDE;  // the grid covers the polygon and a margin around it
    const int c = static_cast<int>(fy) * kCwGrid + static_cast<int>(fx);
    return (B.grid[c >> 3] >> ((c & 7) * 4)) & 0xFu;
  };
  uint32_t all = 0xFu, any = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    const uint32_t c = cell(cx[k], cy[k]);
    all &= c;
    any |= c;
  }
  if (any & CW_CELL_OUTSIDE) return 1;
  if (!(all & CW_CELL_INSIDE)) return -1;
  return (cell(t.x, t.y) & CW_CELL_VERTEX) ? -2 : 0;
}

BRM_HD bool cw_at_goal(const CwBlob& B, const CwKin& t, int a, const double (&cx)[4], const double (&cy)[4]) {
  if (!B.has_goal[a]) return false;
  const double* g = B.goal[a];
  bool in = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) in = in && cx[k] >= g[0] && cx[k] <= g[1] && cy[k] >= g[2] && cy[k] <= g[3];
  return in && t.th >= g[4] && t.th <= g[5];
}

// separating-axis test of the rectangles of agents a and j (centre distance already known to
// be within the sum of the circumscribed radii)
BRM_HD bool cw_rect_overlap(const CwBlob& B, int a, double xa, double ya, double ca, double sa, int j, double xj,
                            double yj, double cj, double sj) {
  const double oi = B.off[a], oj = B.off[j], eix = B.ext[a], ejx = B.ext[j], eiy = B.hw[a], ejy = B.hw[j];
  const double dx = cw_fma(oj, cj, xj) - cw_fma(oi, ca, xa);
  const double dy = cw_fma(oj, sj, yj) - cw_fma(oi, sa, ya);
  const double cd = fabs(cw_fma(ca, cj, sa * sj)), sd = fabs(cw_fma(sa, cj, -(ca * sj)));
  bool hit = fabs(cw_fma(dx, ca, dy * sa)) <= cw_fma(ejy, sd, cw_fma(ejx, cd, eix));
  hit = (fabs(cw_fma(dy, ca, -(dx * sa))) <= cw_fma(ejy, cd, cw_fma(ejx, sd, eiy))) && hit;
  hit = (fabs(cw_fma(dx, cj, dy * sj)) <= cw_fma(eiy, sd, cw_fma(eix, cd, ejx))) && hit;
  hit = (fabs(cw_fma(dy, cj, -(dx * sj))) <= cw_fma(eiy, cd, cw_fma(eix, sd, ejy))) && hit;
  return hit;
}

BRM_HD void cw_load_kin(const double* rec, CwKin& t) {
  t.x = rec[AG_X]; t.y = rec[AG_Y]; t.th = rec[AG_TH]; t.v = rec[AG_V]; t.cs = rec[AG_CS]; t.sn = rec[AG_SN];
}

BRM_HD bool cw_collides(const CwBlob& B, const CwPool& p, int r, int a) {
  const int A = p.A;
  const double* ra = cw_ag(p, cw_slot(p, r, a));
  const double xa = ra[AG_X], ya = ra[AG_Y], ca = ra[AG_CS], sa = ra[AG_SN];
  const double cxa = cw_fma(B.off[a], ca, xa), cya = cw_fma(B.off[a], sa, ya);
  bool coll = false;
  for (int j = 0; j < A; ++j) {
    if (j == a || !(cw_meta(p, cw_slot(p, r, j))->flags & BRM_CW_PRESENT)) continue;
    const double* rj = cw_ag(p, cw_slot(p, r, j));
    const double xj = rj[AG_X], yj = rj[AG_Y], cj = rj[AG_CS], sj = rj[AG_SN];
    const double ddx = cw_fma(B.off[j], cj, xj) - cxa, ddy = cw_fma(B.off[j], sj, yj) - cya;
    const double rr = B.rad[a] + B.rad[j];
    if (cw_fma(ddx, ddx, ddy * ddy) > rr * rr) continue;
    coll = cw_rect_overlap(B, a, xa, ya, ca, sa, j, xj, yj, cj, sj) || coll;
  }
  return coll;
}

// ------------------------------------------------------------------ automaton state packing
BRM_HD uint32_t cw_q_get(const uint32_t* qw, int sl) { return (qw[sl >> 3] >> ((sl & 7) * 4)) & 0xFu; }
BRM_HD void cw_q_set(uint32_t* qw, int sl, uint32_t q) {
  const int sh = (sl & 7) * 4;
  qw[sl >> 3] = (qw[sl >> 3] & ~(0xFu << sh)) | (q << sh);
}
BRM_HD bool cw_slot_absorbing(const CwBlob& B, int sl, uint32_t q) { return (B.slot_rec[sl].z >> (16 + q)) & 1u; }

// ------------------------------------------------------------------ init / finalize items
// what a batch of states looks like to the pool (device or host pointers)
struct CwStatesView {
  const double* agents;  // [A][n][4]
  const uint8_t* flags;  // [A][n]
  const uint8_t* q;      // [M][R][n]
  int64_t n;
};

// agent j of state `leaf` into slot (r, j): pose, lane projection, automaton states
BRM_HD void cw_init_item(const CwBlob& B, CwPool& p, const CwStatesView& S, int r, int j, int64_t leaf) {
  const int M = p.M, R = p.R;
  const int s = cw_slot(p, r, j);
  const double* src = S.agents + (static_cast<int64_t>(j) * S.n + leaf) * 4;
  CwKin t;
  t.x = src[0]; t.y = src[1]; t.th = src[2]; t.v = src[3];
  const uint8_t fl = S.flags[static_cast<int64_t>(j) * S.n + leaf];
  cw_sincos(t.th, &t.sn, &t.cs);
  CwFrenet f;
  f.lane = -1; f.s = 0.0; f.lat = 0.0;
  if (fl & BRM_CW_PRESENT) f = cw_frenet(B, t.x, t.y);
  uint8_t bits = 0;
  if (j == 0 && (fl & BRM_CW_PRESENT) && !(fl & BRM_CW_VALID)) {
    // a frozen ego never passes phase 1: its drivable-area / goal bits are those of the leaf
    double cx[4], cy[4];
    cw_corners(B, t, j, cx, cy);
    if (cw_out_of_map(B, t, j, cx, cy)) bits |= CW_BIT_OOM;
    if (cw_at_goal(B, t, j, cx, cy)) bits |= CW_BIT_GOAL;
  }
  bits |= cw_beyond_bit(B, f);
  double* rec = cw_ag(p, s);
  rec[AG_X] = t.x; rec[AG_Y] = t.y; rec[AG_TH] = t.th; rec[AG_V] = t.v; rec[AG_CS] = t.cs; rec[AG_SN] = t.sn;
  rec[AG_S] = f.s; rec[AG_LAT] = f.lat;
  CwMeta m;
  m.lane = static_cast<int16_t>(f.lane); m.flags = fl; m.bits = bits;
  *cw_meta(p, s) = m;
  rec[AG_COST] = 0.0;
  if (j < M) {
    const int sm = cw_mslot(p, r, j);
    double* mc = cw_mc(p, sm);
    for (int v = 0; v < p.V; ++v) mc[MC_ACC + v] = 0.0;
    uint32_t* rng = cw_mc_rng(p, sm);
    for (int v = 0; v < 4; ++v) rng[v] = 0u;
    uint32_t* qw = cw_mc_qw(p, sm);
    uint8_t* viol = cw_mc_viol(p, sm);
    for (int w = 0; w < p.QW; ++w) qw[w] = 0u;
    uint32_t live = 0u;
    for (int sl = 0; sl < R; ++sl) {
      const uint32_t q = S.q[(static_cast<int64_t>(j) * R + sl) * S.n + leaf];
      cw_q_set(qw, sl, q);
      if (!cw_slot_absorbing(B, sl, q)) live |= 1u << sl;
      viol[sl] = 0;
    }
    *cw_mc_live(p, sm) = live;
  }
}

// GetTerminalReward (:148-164) of MCTS agent a in slot r
BRM_HD void cw_terminal_reward(const CwBlob& B, const CwPool& p, int r, int a, double (&tr)[4]) {
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  if (cw_meta(p, cw_slot(p, r, a))->flags & BRM_CW_PRESENT) {
    const uint32_t* qw = cw_mc_qw(p, cw_mslot(p, r, a));
    for (int sl = 0; sl < p.R; ++sl) {
      const brm_rule& rule = B.rules[B.slot_rule[sl]];
      if (!rule.accepting[cw_q_get(qw, sl)]) {
        const double w = static_cast<double>(rule.weight);
        const int pr = rule.priority;
        t0 = pr == 0 ? t0 + w : t0;
        t1 = pr == 1 ? t1 + w : t1;
        t2 = pr == 2 ? t2 + w : t2;
        t3 = pr == 3 ? t3 + w : t3;
      }
    }
  }
  tr[0] = t0; tr[1] = t1; tr[2] = t2; tr[3] = t3;
}

// ------------------------------------------------------------------ phase 1
struct CwStepCtx {
  uint64_t seed, rollout_base;
  const uint8_t* actions;  // given actions (execute: [M][n], replay: [B][T][M]) or null: Philox
  int64_t n;               // states in the batch (execute action stride)
  int32_t T;               // replay: steps per trace; 0: execute layout
};

BRM_HD uint32_t cw_pick_action(const CwBlob& B, CwPool& p, const CwStepCtx& c, int r, int a) {
  const uint32_t n_prims = B.n_prims[a];
  const CwHdrRef h = cw_hdr(p, r);
  const int k = h[H_K];
  if (c.actions) {
    const int64_t i = h[H_LEAF];
    const uint32_t act = c.T > 0 ? c.actions[(i * c.T + k) * p.M + a] : c.actions[static_cast<int64_t>(a) * c.n + i];
    return act < n_prims ? act : 0u;
  }
  uint32_t* blk = cw_mc_rng(p, cw_mslot(p, r, a));
  if ((k & 3) == 0)
    cw_philox_block(c.seed, c.rollout_base + static_cast<uint64_t>(h[H_FLAT]), static_cast<uint32_t>(k),
                    static_cast<uint32_t>(a), blk);
  return cw_mulhi32(blk[k & 3], n_prims);
}

// Predict for one agent that is present and VALID (:69-71), then what of Execute depends on
// this agent alone: drivable area / goal (:94-96, :110), removal from the map, lane
// projection, GetActionCost (:202-248) and PotentialReward (:250-264).
// Returns true when the drivable-area test could not be decided from the grid: the caller
// then runs cw_exact_shape_item for this agent before phase 2.
BRM_HD bool cw_phase1_item(const CwBlob& B, CwPool& p, const CwStepCtx& c, int r, int j) {
  const int M = p.M;
  const int s = cw_slot(p, r, j);
  double* rec = cw_ag(p, s);
  CwMeta* meta = cw_meta(p, s);
  CwKin t;
  cw_load_kin(rec, t);
  const double v0 = t.v, th0 = t.th;
  const uint32_t act = j < M ? cw_pick_action(B, p, c, r, j) : 0u;
  cw_integrate(B, t, B.prim_a[j][act], B.prim_k[j][act]);
  uint8_t bits = 0;
  bool present = true;
  if (j < M || B.remove_oom) {
    double cx[4], cy[4];
    cw_corners(B, t, j, cx, cy);
    const int cls = cw_classify_shape(B, t, cx, cy);
    if (cls > 0) bits |= CW_BIT_OOM;
    if (cls < 0) bits |= cls == -2 ? (CW_BIT_PENDING | CW_BIT_PENDING_V) : CW_BIT_PENDING;
    if (j < M && cw_at_goal(B, t, j, cx, cy)) bits |= CW_BIT_GOAL;
    if (B.remove_oom && (bits & CW_BIT_OOM)) {
      // World::RemoveInvalidAgents: the agent leaves the world
      present = false;
      bits |= CW_BIT_VANISHED;
    }
  }
  CwFrenet f;
  f.lane = -1; f.s = 0.0; f.lat = 0.0;
  if (present) f = cw_frenet(B, t.x, t.y);
  bits |= cw_beyond_bit(B, f);
  if (j < M) {
    const double dt = B.dt;
    const double a_lon = (t.v - v0) * B.inv_dt;
    double cost = B.w_acc * a_lon * a_lon * dt;
    const double theta_dot = (t.th - th0) * B.inv_dt;
    const double avg_vel = cw_fma(0.5 * a_lon, dt, v0);
    const double a_lat = theta_dot * avg_vel;
    cost = cw_fma(B.w_rad * a_lat * a_lat, dt, cost);
    cost = cw_fma(B.w_vel * fabs(v0 - B.v_des), dt, cost);
    if (f.lane >= 0) cost = cw_fma(B.w_lane * fabs(f.lat), dt, cost);
    const double pot_new = -B.w_pot * dt * fabs(t.v - B.v_des);
    const double pot_cur = -B.w_pot * dt * fabs(v0 - B.v_des);
    rec[AG_COST] = cost + cw_fma(B.gamma, pot_new, -pot_cur);
  }
  rec[AG_X] = t.x; rec[AG_Y] = t.y; rec[AG_TH] = t.th; rec[AG_V] = t.v; rec[AG_CS] = t.cs; rec[AG_SN] = t.sn;
  rec[AG_S] = f.s; rec[AG_LAT] = f.lat;
  CwMeta m;
  m.lane = static_cast<int16_t>(f.lane);
  m.flags = present ? meta->flags : 0;
  m.bits = bits;
  *meta = m;
  return (bits & CW_BIT_PENDING) != 0;
}

// the exact drivable-area test for an agent phase 1 left pending (near the polygon's boundary)
BRM_HD void cw_exact_shape_item(const CwBlob& B, CwPool& p, int r, int j) {
  const int s = cw_slot(p, r, j);
  double* rec = cw_ag(p, s);
  CwMeta* meta = cw_meta(p, s);
  CwKin t;
  cw_load_kin(rec, t);
  double cx[4], cy[4];
  cw_corners(B, t, j, cx, cy);
  CwMeta m = *meta;
  const bool vertex_only = (m.bits & CW_BIT_PENDING_V) != 0;
  m.bits &= static_cast<uint8_t>(~(CW_BIT_PENDING | CW_BIT_PENDING_V));
  if (vertex_only ? cw_vertex_in_shape(B, t, j) : cw_out_of_map(B, t, j, cx, cy)) {
    m.bits |= CW_BIT_OOM;
    if (B.remove_oom) {
      m.bits |= CW_BIT_VANISHED;
      m.bits &= static_cast<uint8_t>(~CW_BIT_BEYOND);
      m.flags = 0;
      m.lane = -1;
      rec[AG_S] = 0.0;
      rec[AG_LAT] = 0.0;
    }
  }
  *meta = m;
}

// drivable-area / goal bits of MCTS agent a in slot r as it stands (EvaluateRules() on a state
// that is not the result of a step in this pool)
BRM_HD void cw_shape_bits_item(const CwBlob& B, CwPool& p, int r, int a) {
  const int s = cw_slot(p, r, a);
  CwKin t;
  cw_load_kin(cw_ag(p, s), t);
  double cx[4], cy[4];
  cw_corners(B, t, a, cx, cy);
  uint8_t bits = cw_meta(p, s)->bits & CW_BIT_BEYOND;
  if (cw_out_of_map(B, t, a, cx, cy)) bits |= CW_BIT_OOM;
  if (cw_at_goal(B, t, a, cx, cy)) bits |= CW_BIT_GOAL;
  cw_meta(p, s)->bits = bits;
}

// ------------------------------------------------------------------ phase 2
struct CwEval {
  double r[4];
  uint32_t w[4];  // labels: ego bits by kind, masks over j for kinds 8, 9, 10
};

// Everything of Execute that needs the other agents (:87-120), for MCTS agent a of slot r:
// collision, labels, rule automata, reward vector, validity; the ego decides CheckTerminal
// (:266-285) of the new state.  `rules_only`: MvmctsState::EvaluateRules() (:186-197) on the
// state as it is.  Returns through `o`; accumulates disc * reward into the slot.
BRM_HD void cw_phase2_item(const CwBlob& B, CwPool& p, int r, int a, bool rules_only, CwEval& o) {
  const int A = p.A, V = p.V;
  const int s = cw_slot(p, r, a), sm = cw_mslot(p, r, a);
  o.r[0] = o.r[1] = o.r[2] = o.r[3] = 0.0;
  o.w[0] = o.w[1] = o.w[2] = o.w[3] = 0u;
  CwMeta* meta = cw_meta(p, s);
  const CwMeta mt = *meta;
  const uint8_t fl = mt.flags, bits = mt.bits;
  double* mc = cw_mc(p, sm);
  const bool present = fl & BRM_CW_PRESENT;
  if (!present) {
    if ((bits & CW_BIT_VANISHED) && !rules_only) {
      // left the map during this step: the out-of-map penalty and nothing else (:121-125)
      o.r[0] = B.w_oom;
      mc[MC_ACC] = cw_fma(p.disc[r], o.r[0], mc[MC_ACC]);
      meta->bits = 0;
      if (a == 0) cw_hstate(p, r)[HS_DONE] = 1;  // CheckTerminal: no ego
    }
    return;
  }
  const bool evaluated = (fl & BRM_CW_VALID) != 0;
  const bool oom = bits & CW_BIT_OOM, goal = bits & CW_BIT_GOAL;
  // pairwise predicates against the other agents of the rollout
  const double* ra = cw_ag(p, s);
  const double xa = ra[AG_X], ya = ra[AG_Y], ca = ra[AG_CS], sa = ra[AG_SN], va = ra[AG_V], s_a = ra[AG_S];
  const int lane_a = mt.lane;
  const double cxa = cw_fma(B.off[a], ca, xa), cya = cw_fma(B.off[a], sa, ya);
  const double my_len = lane_a >= 0 ? B.lanes[lane_a].length : 0.0;
  const int my_succ = lane_a >= 0 ? B.lanes[lane_a].succ : -2;
  bool coll = false;
  int jf = -1;
  double gf = 0.0, vf = 0.0;
  uint32_t merged_mask = 0, near_mask = 0;
  // (the other agents in ascending order, skipping myself without a divergent iteration: the lanes
  // of a warp are different agents, and `if (j == a) continue` would idle a quarter of them in turn)
  for (int o = 0; o < A - 1; ++o) {
    const int j = o + (o >= a ? 1 : 0);
    const int sj = cw_slot(p, r, j);
    const CwMeta mj = *cw_meta(p, sj);
    if (!(mj.flags & BRM_CW_PRESENT)) continue;
    const double* rj = cw_ag(p, sj);
    const double xj = rj[AG_X], yj = rj[AG_Y];
    const double cj = rj[AG_CS], snj = rj[AG_SN];
    const double ddx = cw_fma(B.off[j], cj, xj) - cxa, ddy = cw_fma(B.off[j], snj, yj) - cya;
    const double rr = B.rad[a] + B.rad[j];
    if (cw_fma(ddx, ddx, ddy * ddy) <= rr * rr) coll = cw_rect_overlap(B, a, xa, ya, ca, sa, j, xj, yj, cj, snj) || coll;
    if (!evaluated) continue;
    const int lane_j = mj.lane;
    if (lane_j >= 0) {
      const double s_j = rj[AG_S];
      if (lane_a >= 0) {
        // gap along my corridor: own lane, then its successor
        double g = 0.0;
        bool has = false;
        if (lane_j == lane_a) {
          g = s_j - s_a;
          has = true;
        } else if (my_succ == lane_j) {
          g = (my_len - s_a) + s_j;
          has = true;
        }
        if (has && g > 0.0 && (jf < 0 || g < gf)) {
          jf = j;
          gf = g;
          vf = rj[AG_V];
        }
      }
    }
    if (mj.bits & CW_BIT_BEYOND) merged_mask |= 1u << j;
    const double dx = xj - xa, dy = yj - ya;
    if (cw_fma(dx, dx, dy * dy) < B.near2) near_mask |= 1u << j;
  }
  if (evaluated) {
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
    auto add_at = [&](int idx, double val) {
      r0 = idx == 0 ? r0 + val : r0;
      r1 = idx == 1 ? r1 + val : r1;
      r2 = idx == 2 ? r2 + val : r2;
      r3 = idx == 3 ? r3 + val : r3;
    };
    if (!rules_only) {
      r0 += (coll ? 1.0 : 0.0) * B.w_coll;  // :91-92
      r0 += (oom ? 1.0 : 0.0) * B.w_oom;    // :94-96
      add_at(V - 1, cw_ag(p, s)[AG_COST]);  // :98-102
    }
    // labels (EvaluateRules, :165-185)
    uint32_t w0 = 0;
    if (coll) w0 |= 1u << BRM_CW_LABEL_COLLISION_EGO;
    if (oom) w0 |= 1u << BRM_CW_LABEL_COLLISION_CORRIDOR;
    if (goal) w0 |= 1u << BRM_CW_LABEL_GOAL_REACHED;
    bool safe = true;
    if (jf >= 0) {
      const double dist = gf - (B.front[a] + B.rear[jf]);
      const double dsafe = cw_fma(va, B.sd_delta, (va * va) * B.sd_ar) - (vf * vf) * B.sd_af;
      safe = dist >= dsafe;
    }
    if (safe) w0 |= 1u << BRM_CW_LABEL_SAFE_DISTANCE_FRONT;
    if (lane_a >= 0) {
      const CwLaneHdr& L = B.lanes[lane_a];
      if (bits & CW_BIT_BEYOND) w0 |= 1u << BRM_CW_LABEL_MERGED_E;
      if (L.flags & 1) w0 |= 1u << BRM_CW_LABEL_RIGHTMOST_LANE;
      w0 |= 1u << BRM_CW_LABEL_ON_ROAD;
    }
    const uint32_t front_mask = jf >= 0 ? (1u << jf) : 0u;
    o.w[0] = w0; o.w[1] = merged_mask; o.w[2] = front_mask; o.w[3] = near_mask;
    if (rules_only || !(B.ego_only && a != 0)) {  // :104-108
      uint32_t* qw = cw_mc_qw(p, sm);
      uint8_t* viol = cw_mc_viol(p, sm);
      uint32_t live = *cw_mc_live(p, sm);
      uint32_t todo = live;
      double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
      while (todo) {
#if defined(__CUDA_ARCH__)
        const int sl = __ffs(static_cast<int>(todo)) - 1;
#else
        const int sl = __builtin_ctz(todo);
#endif
        todo &= todo - 1u;
        const CwSlotRec rec = B.slot_rec[sl];
        const int rank = static_cast<int>((rec.z >> 8) & 0xffu) - 1;
        const int other = rank < 0 ? 0 : (rank < a ? rank : rank + 1);
        const uint32_t n_aps = rec.y & 0xffu;
        const uint32_t cur = cw_q_get(qw, sl);
        // label word of this slot: ego kinds in bits 0..7, the bound agent's kinds in 8..10
        const uint32_t W = w0 | (((merged_mask >> other) & 1u) << BRM_CW_LABEL_MERGED_X) |
                           (((front_mask >> other) & 1u) << BRM_CW_LABEL_IN_DIRECT_FRONT_X) |
                           (((near_mask >> other) & 1u) << BRM_CW_LABEL_NEAR_X);
        uint32_t letter = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t k = 0; k < BRM_MAX_APS; ++k) letter |= ((W >> ((rec.x >> (8 * k)) & 31u)) & 1u) << k;
        uint32_t nxt = B.rules[rec.z & 0xffu].trans[(cur << n_aps) + letter];
        if (nxt == BRM_VIOLATION) {
          viol[sl] += 1;
          nxt = (rec.y >> 24) ? cur : ((rec.y >> 16) & 0xffu);
          const int pr = static_cast<int>((rec.y >> 8) & 0xffu);
          const double wgt = static_cast<double>(rec.w);
          q0 = pr == 0 ? q0 + wgt : q0;
          q1 = pr == 1 ? q1 + wgt : q1;
          q2 = pr == 2 ? q2 + wgt : q2;
          q3 = pr == 3 ? q3 + wgt : q3;
        }
        if (nxt != cur) {
          cw_q_set(qw, sl, nxt);
          if (cw_slot_absorbing(B, sl, nxt)) live &= ~(1u << sl);
        }
      }
      *cw_mc_live(p, sm) = live;
      r0 += q0; r1 += q1; r2 += q2; r3 += q3;
    }
    if (goal && !rules_only) add_at(V - 1, B.goal_reward);                                 // :110-113
    if (coll && !rules_only) meta->flags = fl & static_cast<uint8_t>(~BRM_CW_VALID);       // :116-120
    o.r[0] = r0; o.r[1] = r1; o.r[2] = r2; o.r[3] = r3;
    const double disc = p.disc[r];
    mc[MC_ACC + 0] = cw_fma(disc, r0, mc[MC_ACC + 0]);  // the record holds V accumulators
    if (V > 1) mc[MC_ACC + 1] = cw_fma(disc, r1, mc[MC_ACC + 1]);
    if (V > 2) mc[MC_ACC + 2] = cw_fma(disc, r2, mc[MC_ACC + 2]);
    if (V > 3) mc[MC_ACC + 3] = cw_fma(disc, r3, mc[MC_ACC + 3]);
  }
  // CheckTerminal of the new state (:266-285)
  if (a == 0 && !rules_only && (cw_hdr(p, r)[H_HORIZON] - 1 == 0 || oom || coll || goal)) cw_hstate(p, r)[HS_DONE] = 1;
}

// MvmctsState::CheckTerminal (:266-285) of the state in slot r as it is
BRM_HD bool cw_check_terminal(const CwBlob& B, const CwPool& p, int r) {
  const int s = cw_slot(p, r, 0);
  if (!(cw_meta(p, s)->flags & BRM_CW_PRESENT)) return true;
  CwKin t;
  cw_load_kin(cw_ag(p, s), t);
  double cx[4], cy[4];
  cw_corners(B, t, 0, cx, cy);
  return cw_hdr(p, r)[H_HORIZON] == 0 || cw_out_of_map(B, t, 0, cx, cy) || cw_collides(B, p, r, 0) ||
         cw_at_goal(B, t, 0, cx, cy);
}

}  // namespace brm
