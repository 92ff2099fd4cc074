// TEST INFRASTRUCTURE -- runs the device arithmetic of the continuous-world kernels
// (planner-rules-mcts_b200/csrc/cw_core.cuh: the work-item functions the CUDA kernels are made
// of) on the CPU, item by item in the order the kernels' phases impose, so that the step
// arithmetic can be compared with the oracle without a GPU.  cw_core.cuh only uses fma/+/*
// and the library is built with --fmad=false, so this is the arithmetic of the GPU, bit for
// bit.  Not part of the product: libbrm.so contains none of this file, and nothing under
// planner-rules-mcts_b200/ loads it.
//   g++ -std=c++17 -O2 -ffp-contract=off -mfma -shared -fPIC -I include -I planner-rules-mcts_b200/csrc
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cw_blob_build.h"
#include "cw_core.cuh"

using namespace brm;

namespace {

struct Emul {
  CwBlob* B = nullptr;
  int R = 0;
  CwPool pool;
  void* mem = nullptr;
  std::string err;
  ~Emul() {
    free(B);
    free(mem);
  }
  bool setup(const brm_cw_scenario* scn, int P) {
    B = static_cast<CwBlob*>(aligned_alloc(16, sizeof(CwBlob)));
    if (!cw_build_blob(*scn, B, &R, &err)) return false;
    const size_t bytes = cw_pool_bytes(P, B->A, B->M, R, B->V);
    mem = aligned_alloc(16, bytes);
    memset(mem, 0xA5, bytes);  // stale contents must never matter
    cw_pool_carve(pool, mem, P, B->A, B->M, R, B->V);
    for (int r = 0; r < P; ++r) cw_hstate(pool, r)[HS_BUSY] = 0;
    return true;
  }
};

thread_local std::string g_err;

// one state as a batch of 1
struct OneState {
  std::vector<double> agents;
  std::vector<uint8_t> flags, q;
  CwStatesView view(int A) {
    CwStatesView v;
    v.agents = agents.data();
    v.flags = flags.data();
    v.q = q.data();
    v.n = 1;
    (void)A;
    return v;
  }
};

OneState make_state(const CwBlob& B, int R, const double* states0, const uint8_t* flags0, const uint8_t* q0) {
  OneState s;
  s.agents.assign(states0, states0 + B.A * 4);
  s.flags.assign(flags0, flags0 + B.A);
  s.q.resize(static_cast<size_t>(B.M) * R);
  for (int a = 0; a < B.M; ++a)
    for (int k = 0; k < R; ++k)
      s.q[a * R + k] = q0 ? q0[a * R + k] : B.rules[B.slot_rule[k]].init_state;
  return s;
}

void fill_slot(Emul& e, OneState& st, int r, int horizon, int flat, double disc0) {
  CwPool& p = e.pool;
  const CwStatesView v = st.view(p.A);
  const CwHdrRef h = cw_hdr(p, r);
  h[H_FLAT] = flat;
  h[H_LEAF] = 0;
  h[H_K] = 0;
  h[H_HORIZON] = horizon;
  h[H_NSTEPS] = 0;
  p.disc[r] = disc0;
  cw_hstate(p, r)[HS_BUSY] = 1;
  cw_hstate(p, r)[HS_DONE] = 0;
  for (int j = 0; j < p.A; ++j) cw_init_item(*e.B, p, v, r, j, 0);
}

// one Execute on slot r: the kernels' list building, phase 1, phase 2 and bookkeeping
void step_slot(Emul& e, const CwStepCtx& c, int r, CwEval* ev /*[M]*/, double gamma) {
  CwPool& p = e.pool;
  const CwBlob& B = *e.B;
  const int A = p.A, M = p.M;
  std::vector<int> l1, l2;
  for (int j = 0; j < A; ++j) {
    const uint8_t f = cw_meta(p, cw_slot(p, r, j))->flags;
    if ((f & BRM_CW_PRESENT) && (f & BRM_CW_VALID)) l1.push_back(j);
    if (j < M && (f & BRM_CW_PRESENT) && ((f & BRM_CW_VALID) || j == 0)) l2.push_back(j);
  }
  std::vector<int> lx;
  for (int j : l1) {
    if (cw_phase1_item(B, p, c, r, j)) lx.push_back(j);
    cw_hdr(p, r)[H_NSTEPS] += 1;
  }
  for (int j : lx) cw_exact_shape_item(B, p, r, j);  // the kernels' phase 1b
  for (int a = 0; a < M; ++a) {
    memset(&ev[a], 0, sizeof(CwEval));
  }
  for (int a : l2) cw_phase2_item(B, p, r, a, false, ev[a]);
  cw_hdr(p, r)[H_K] += 1;
  p.disc[r] *= gamma;
  cw_hdr(p, r)[H_HORIZON] -= 1;
}

}  // namespace

extern "C" {

const char* cwe_last_error() { return g_err.c_str(); }

void cwe_sincos(double x, double* s, double* c) { cw_sincos(x, s, c); }

int cwe_frenet(const brm_cw_scenario* scn, double x, double y, double* out /*lane, s, lat*/) {
  Emul e;
  if (!e.setup(scn, 1)) {
    g_err = e.err;
    return -1;
  }
  const CwFrenet f = cw_frenet(*e.B, x, y);
  out[0] = f.lane;
  out[1] = f.s;
  out[2] = f.lat;
  return 0;
}

// fraction of phase-1 items whose drivable-area test the grid could not decide (diagnostic)
int cwe_classify(const brm_cw_scenario* scn, double x, double y, double th, int agent) {
  Emul e;
  if (!e.setup(scn, 1)) {
    g_err = e.err;
    return -2;
  }
  CwKin t;
  t.x = x; t.y = y; t.th = th; t.v = 0.0;
  cw_sincos(th, &t.sn, &t.cs);
  double cx[4], cy[4];
  cw_corners(*e.B, t, agent, cx, cy);
  const int cls = cw_classify_shape(*e.B, t, cx, cy);
  const int exact = cw_out_of_map(*e.B, t, agent, cx, cy) ? 1 : 0;
  // "corners certainly inside": the vertex test alone must then say what the exact test says
  if (cls == -2 && (cw_vertex_in_shape(*e.B, t, agent) ? 1 : 0) != exact) return -5;
  return cls < 0 ? 2 + exact : (cls == exact ? cls : -3 - cls);  // 0/1 decided and right, 2/3 pending, < -2 WRONG
}

int cwe_check_terminal(const brm_cw_scenario* scn, const double* states0, const uint8_t* flags0, int horizon0) {
  Emul e;
  if (!e.setup(scn, 2)) {
    g_err = e.err;
    return -1;
  }
  OneState st = make_state(*e.B, e.R, states0, flags0, nullptr);
  fill_slot(e, st, 1, horizon0, 0, 1.0);
  return cw_check_terminal(*e.B, e.pool, 1) ? 1 : 0;
}

// same outputs as cwo_replay (oracle/cw_oracle.h) except the margins
int cwe_replay(const brm_cw_scenario* scn, const double* states0, const uint8_t* flags0, int horizon0,
               const uint8_t* q0, const uint8_t* actions, int T, double* states_out, uint8_t* flags_out,
               uint8_t* terminal_out, uint8_t* q_out, uint32_t* viol_out, double* rewards_out,
               uint32_t* labels_out, double* frenet_out, double* term_reward_out) {
  Emul e;
  if (!e.setup(scn, 3)) {
    g_err = e.err;
    return -1;
  }
  const CwBlob& B = *e.B;
  CwPool& p = e.pool;
  const int A = B.A, M = B.M, V = B.V, R = e.R, r = 2;
  OneState st = make_state(B, R, states0, flags0, q0);
  fill_slot(e, st, r, horizon0, 0, 1.0);
  bool terminal = cw_check_terminal(B, p, r);
  CwStepCtx c;
  c.seed = 0; c.rollout_base = 0; c.actions = actions; c.n = 1; c.T = T;
  std::vector<CwEval> ev(M);
  std::vector<uint32_t> viol32(static_cast<size_t>(M) * R, 0);
  int t = 0;
  for (; t < T && !terminal; ++t) {
    cw_hstate(p, r)[HS_DONE] = 0;
    step_slot(e, c, r, ev.data(), 1.0);
    terminal = cw_hstate(p, r)[HS_DONE] != 0;
    for (int i = 0; i < A; ++i) {
      const int s = cw_slot(p, r, i);
      const double* rec = cw_ag(p, s);
      double* o = states_out + (static_cast<size_t>(t) * A + i) * 4;
      o[0] = rec[AG_X]; o[1] = rec[AG_Y]; o[2] = rec[AG_TH]; o[3] = rec[AG_V];
      flags_out[t * A + i] = cw_meta(p, s)->flags;
      double* fo = frenet_out + (static_cast<size_t>(t) * A + i) * 3;
      fo[0] = cw_meta(p, s)->lane; fo[1] = rec[AG_S]; fo[2] = rec[AG_LAT];
    }
    terminal_out[t] = terminal;
    for (int a = 0; a < M; ++a) {
      const int sm = cw_mslot(p, r, a);
      for (int k = 0; k < R; ++k) {
        viol32[a * R + k] += cw_mc_viol(p, sm)[k];
        cw_mc_viol(p, sm)[k] = 0;
        q_out[(static_cast<size_t>(t) * M + a) * R + k] = static_cast<uint8_t>(cw_q_get(cw_mc_qw(p, sm), k));
        viol_out[(static_cast<size_t>(t) * M + a) * R + k] = viol32[a * R + k];
      }
      for (int v = 0; v < V; ++v) rewards_out[(static_cast<size_t>(t) * M + a) * V + v] = ev[a].r[v];
      for (int w = 0; w < 4; ++w) labels_out[(static_cast<size_t>(t) * M + a) * 4 + w] = ev[a].w[w];
    }
  }
  for (int a = 0; a < M; ++a) {
    double tr[4];
    cw_terminal_reward(B, p, r, a, tr);
    for (int v = 0; v < V; ++v) term_reward_out[a * V + v] = tr[v];
  }
  return t;
}

// same outputs as cwo_rollout_detail except the margins
int cwe_rollout_detail(const brm_cw_scenario* scn, const double* states0, const uint8_t* flags0, int horizon0,
                       const uint8_t* q0, uint64_t seed, uint64_t rollout_begin, uint64_t n_rollouts, int max_steps,
                       double gamma, int first_exp, int leaf_terminal, double* returns, uint32_t* viol,
                       uint8_t* q_final, double* states_final, uint8_t* flags_final, int32_t* steps,
                       uint64_t* agent_steps) {
  const int P = 5;
  Emul e;
  if (!e.setup(scn, P)) {
    g_err = e.err;
    return -1;
  }
  const CwBlob& B = *e.B;
  CwPool& p = e.pool;
  const int A = B.A, M = B.M, V = B.V, R = e.R;
  OneState st = make_state(B, R, states0, flags0, q0);
  CwStepCtx c;
  c.seed = seed; c.rollout_base = rollout_begin; c.actions = nullptr; c.n = 1; c.T = 0;
  std::vector<CwEval> ev(M);
  *agent_steps = 0;
  for (uint64_t i = 0; i < n_rollouts; ++i) {
    const int r = static_cast<int>(i % P);  // slots are reused without being cleared
    fill_slot(e, st, r, horizon0, static_cast<int>(i), first_exp ? gamma : 1.0);
    cw_hstate(p, r)[HS_DONE] = leaf_terminal ? 1 : 0;
    while (!cw_hstate(p, r)[HS_DONE] && cw_hdr(p, r)[H_K] < max_steps) step_slot(e, c, r, ev.data(), gamma);
    for (int a = 0; a < M; ++a) {
      const int sm = cw_mslot(p, r, a);
      double tr[4];
      cw_terminal_reward(B, p, r, a, tr);
      for (int v = 0; v < V; ++v) returns[(i * M + a) * V + v] = fma(p.disc[r], tr[v], cw_mc(p, sm)[MC_ACC + v]);
      for (int k = 0; k < R; ++k) {
        viol[(i * M + a) * R + k] = cw_mc_viol(p, sm)[k];
        q_final[(i * M + a) * R + k] = static_cast<uint8_t>(cw_q_get(cw_mc_qw(p, sm), k));
      }
    }
    for (int j = 0; j < A; ++j) {
      const int s = cw_slot(p, r, j);
      const double* rec = cw_ag(p, s);
      double* o = states_final + (i * A + j) * 4;
      o[0] = rec[AG_X]; o[1] = rec[AG_Y]; o[2] = rec[AG_TH]; o[3] = rec[AG_V];
      flags_final[i * A + j] = cw_meta(p, s)->flags;
    }
    steps[i] = cw_hdr(p, r)[H_K];
    *agent_steps += static_cast<uint32_t>(cw_hdr(p, r)[H_NSTEPS]);
  }
  return 0;
}

// MvmctsState::EvaluateRules() on one state
int cwe_evaluate_rules(const brm_cw_scenario* scn, const double* states0, const uint8_t* flags0, int horizon0,
                       const uint8_t* q0, uint8_t* q_out, uint32_t* viol_out, double* rewards_out,
                       uint32_t* labels_out) {
  Emul e;
  if (!e.setup(scn, 1)) {
    g_err = e.err;
    return -1;
  }
  const CwBlob& B = *e.B;
  CwPool& p = e.pool;
  const int M = B.M, V = B.V, R = e.R;
  OneState st = make_state(B, R, states0, flags0, q0);
  fill_slot(e, st, 0, horizon0, 0, 1.0);
  for (int a = 0; a < M; ++a) {
    CwEval ev;
    memset(&ev, 0, sizeof(ev));
    const uint8_t f = cw_meta(p, a)->flags;
    if ((f & BRM_CW_PRESENT) && (f & BRM_CW_VALID)) {
      cw_shape_bits_item(B, p, 0, a);  // the drivable-area / goal labels of the state as it is
      cw_phase2_item(B, p, 0, a, true, ev);
    }
    for (int k = 0; k < R; ++k) {
      q_out[a * R + k] = static_cast<uint8_t>(cw_q_get(cw_mc_qw(p, a), k));
      viol_out[a * R + k] = cw_mc_viol(p, a)[k];
    }
    for (int v = 0; v < V; ++v) rewards_out[a * V + v] = ev.r[v];
    for (int w = 0; w < 4; ++w) labels_out[a * 4 + w] = ev.w[w];
  }
  return 0;
}

}  // extern "C"
