#!/usr/bin/env python
"""Benchmark of the rollout hot path (BASELINE.json metric: rollout-steps/sec, agent-steps,
device-timed).

One "step" = one decision's worth of rollouts: `leaves` leaf states of a host-side search
tree x `rollouts_per_leaf` random rollouts each (mvmcts::RandomHeuristic), fused on the GPU
(K3 rules_rollout / K1 gridworld_rollout), reduced per leaf, folded into root statistics
(K5) and -- with more than one rank -- merged by one NCCL allreduce.  Workloads:
  M  multi-agent merge, 4 agents, zipper + safe-distance rules, 262144 rollouts/decision
     (BASELINE.json configs[2]; the configuration north_star quotes its target on) [default]
  S  single-agent highway, safe-distance rule, 65536 rollouts/decision, horizon 20 (configs[1])
  G  GridWorld threshold-test scenario, 262144 rollouts/decision (configs[0] shape)
  W  4096-intersection sweep shape, 10 agents x 5 rules, 64 rollouts/scenario (configs[4])

`python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
`--impl reference` times the CPU implementation of the same path on the host cores
(oracle/_ref = the reference's own GridWorld code for workload G, the CPU restatement in
oracle/ for S/M/W -- the only place besides cpu_baseline where bench.py executes oracle/).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rollout-steps/sec (agent-steps, device-timed)"
UNIT = "agent-steps/s"
L2_FLUSH_BYTES = 256 << 20

WORKLOADS = {
    "M": dict(name="multi_agent_merge", leaves=256, rollouts_per_leaf=1024, max_steps=20, agents=4,
              desc="multi-agent merge: 4 agents, zipper + safe-distance rules, 262144 rollouts/decision, horizon 20"),
    "R": dict(name="root_parallel_merge", leaves=128, rollouts_per_leaf=1024, max_steps=20, agents=4,
              desc="root-parallel MCTS: the merge scenario with 131072 rollouts/decision per GPU (1M on 8 GPUs), "
                   "per-GPU root statistics merged by one NCCL allreduce"),
    "S": dict(name="single_agent_highway", leaves=64, rollouts_per_leaf=1024, max_steps=20, agents=4,
              desc="single-agent highway lane change, safe-distance rule, 65536 rollouts/decision, horizon 20"),
    "G": dict(name="gridworld_threshold_test", leaves=256, rollouts_per_leaf=1024, max_steps=30, agents=3,
              desc="GridWorld BaseTestEnv: 3 agents, G !collision + ZIP rules, 262144 rollouts/decision, depth 8"),
    "W": dict(name="intersection_sweep", leaves=512, rollouts_per_leaf=64, max_steps=20, agents=10,
              desc="scenario sweep: 512 synthetic intersections per GPU x 10 agents x 5 rules, 64 rollouts each"),
}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (the timed region); if the region was too
        short to catch one, of all samples since start() (warm-up + timed region)."""
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1)]
        if not rows:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than one sample)"
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), window=window)


# ----------------------------------------------------------------------------- workloads
def first_actions(kind, rank, L, M, n_act):
    """The root joint action every leaf of the decision descends from (same on both arms)."""
    if kind == "W":
        return np.zeros((L, M), np.uint8)
    return np.random.default_rng(1000 + rank).integers(0, n_act, size=(L, M)).astype(np.uint8)


def workload_scenarios(kind, rank, L):
    from planner_rules_mcts_b200 import scenarios
    if kind == "W":
        return [scenarios.intersection_scenario(seed=1000, scenario_id=rank * L + k) for k in range(L)]
    return [scenarios.merge_scenario(seed=1000) if kind in ("M", "R") else scenarios.highway_scenario(seed=1000)]


def build_workload(kind, device, rank, cuda_stream=None):
    """Returns a dict with the engine, host leaves, first actions, sizes and oracle hooks."""
    import planner_rules_mcts_b200 as brm
    w = dict(WORKLOADS[kind])
    L = w["leaves"]
    if kind == "G":
        eng = brm.GridWorldEngine(device=device, cuda_stream=cuda_stream)
        root = eng.base_test_env_batch(L)
        first = first_actions(kind, rank, L, eng.A, 3)
        leaves, _ = eng.execute(root, first.T.copy())
        leaves.depth[:] = 0  # TestRunner::RunTest resets the depth after every real step
        w.update(engine=eng, leaves=leaves, first=first, M=eng.A, A=eng.A, V=eng.V, R=eng.R, n_actions=3,
                 gamma=0.95, dtype="i32", state_bytes=16, scn=None)
        return w
    scns = workload_scenarios(kind, rank, L)
    eng = brm.RulesMctsEngine(scns, device=device, cuda_stream=cuda_stream)
    if kind == "W":
        leaves = eng.initial_batch()
        first = first_actions(kind, rank, L, eng.M, 3)
        w.update(engine=eng, leaves=leaves, first=first, M=eng.M, A=eng.A, V=eng.V, R=eng.R, n_actions=3,
                 gamma=0.9, dtype="f64", state_bytes=32, scn=scns)
        return w
    scn = scns[0]
    n_act = len(scn.agents[0].primitives)
    # leaves = children of the root under random first joint actions (tree kept on the host)
    root = eng.make_batch(np.broadcast_to(scn.initial_state, (L,) + scn.initial_state.shape))
    first = first_actions(kind, rank, L, eng.M, n_act)
    leaves, _ = eng.execute(root, first.T.copy())
    assert not leaves.terminal.any(), "a child of the root is terminal: pick another seed"
    w.update(engine=eng, leaves=leaves, first=first, M=eng.M, A=eng.A, V=eng.V, R=eng.R, n_actions=n_act,
             gamma=scn.params.DISCOUNT_FACTOR, dtype="f64", state_bytes=32, scn=[scn])
    return w


class _HostLeaves:
    pass


def reference_workload(kind, rank=0):
    """The SAME leaves as the GPU arm of this rank, produced without a GPU: the root's children
    under the same first joint actions, stepped by the CPU implementation itself (one replayed
    Execute per leaf).  Nothing of libbrm.so is loaded on this path."""
    w = dict(WORKLOADS[kind])
    L = w["leaves"]
    if kind == "G":
        from oracle import gw as ogw
        import planner_rules_mcts_b200.gridworld as gwm   # host-side rule tables only (no engine)
        from tests.gw_common import oracle_config_from_engine
        have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgw_ref.so"))
        orc = ogw.GwOracle("ref" if have_ref else "port")

        class _E:  # what oracle_config_from_engine reads
            params = gwm.GridWorldStateParameter()
            rules = gwm.make_default_rules()
        cfg = oracle_config_from_engine(_E)
        agents0, term0 = ogw.base_test_env_state()
        first = first_actions(kind, rank, L, 3, 3)
        R = orc.rules_per_agent(cfg)
        lv = _HostLeaves()
        lv.agents = np.zeros((3, L, 4), np.int32)
        lv.term = np.zeros(L, np.uint8)
        lv.q = np.zeros((3, R, L), np.uint8)
        for l in range(L):
            o = orc.replay(cfg, agents0, term0, 0, first[l][None, :])
            lv.agents[:, l, :] = o["agents"][0]
            lv.term[l] = sum(int(o["term"][0, a]) << a for a in range(3))
            lv.q[:, :, l] = o["q"][0]
        w.update(lv=lv, n_leaves=L, cfg=cfg, orc=orc, A=3, gamma=0.95, kind_name="reference" if have_ref else "port")
        return w
    from oracle import cw as ocw
    from tests.cw_common import oracle_config
    scns = workload_scenarios(kind, rank, L)
    orc = ocw.CwOracle()
    A, M = len(scns[0].agents), scns[0].n_mcts_agents
    n_act = len(scns[0].agents[0].primitives)
    first = first_actions(kind, rank, L, M, n_act)
    lv = _HostLeaves()
    lv.scenario = np.arange(L, dtype=np.int32) if kind == "W" else np.zeros(L, np.int32)
    lv.agents = np.zeros((A, L, 4))
    lv.flags = np.zeros((A, L), np.uint8)
    lv.horizon = np.zeros(L, np.int32)
    cfgs = [oracle_config(s) for s in scns]
    R = orc.rules_per_agent(cfgs[0])
    lv.q = np.zeros((M, R, L), np.uint8)
    for l in range(L):
        scn, cfg = scns[lv.scenario[l]], cfgs[lv.scenario[l]]
        fl = np.full(A, 3, np.uint8)
        if kind == "W":   # the sweep starts every scenario from its root
            lv.agents[:, l, :], lv.flags[:, l], lv.horizon[l] = scn.initial_state, fl, scn.params.HORIZON
            continue
        o = orc.replay(cfg, scn.initial_state, fl, scn.params.HORIZON, first[l][None, :])
        assert o["steps"] == 1 and not o["terminal"][0]
        lv.agents[:, l, :], lv.flags[:, l], lv.horizon[l] = o["states"][0], o["flags"][0], scn.params.HORIZON - 1
        lv.q[:, :, l] = o["q"][0]
    if kind == "W":
        s = 0
        for r in scns[0].rules:      # initial automaton states (host-side table data)
            reps = (A - 1) if r.IsAgentSpecific() else 1
            lv.q[:, s:s + reps, :] = r.table.init
            s += reps
    w.update(lv=lv, n_leaves=L, cfgs=cfgs, scn=scns, orc=orc, A=A, gamma=scns[0].params.DISCOUNT_FACTOR, kind_name="port")
    return w


def cpu_rollouts(kind, w, leaf_ids, n_per_leaf, n_threads, seed, rollout_base):
    """Rollouts [rollout_base + l * rollouts_per_leaf, ... + n_per_leaf) of the leaves `leaf_ids` on
    `n_threads` host threads -- the ids the GPU arm gives the same leaves.  One leaf per task, the
    tasks spread over a pool of `n_threads` workers (the oracle call releases the GIL), so every
    core runs whole leaves back to back.  Returns (agent_steps, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    orc, lv = w["orc"], w["lv"]
    n = w["rollouts_per_leaf"]

    def one(l):
        base = rollout_base + l * n
        if kind == "G":
            term0 = np.array([(int(lv.term[l]) >> a) & 1 for a in range(w["A"])], np.uint8)
            r = orc.rollout(w["cfg"], np.ascontiguousarray(lv.agents[:, l, :]), term0, 0, seed, base, n_per_leaf,
                            w["max_steps"], w["gamma"], n_threads=1, q0=np.ascontiguousarray(lv.q[:, :, l]))
        else:
            r = orc.rollout(w["cfgs"][int(lv.scenario[l])], np.ascontiguousarray(lv.agents[:, l, :]),
                            np.ascontiguousarray(lv.flags[:, l]), int(lv.horizon[l]), seed, base, n_per_leaf,
                            w["max_steps"], w["gamma"], n_threads=1, q0=np.ascontiguousarray(lv.q[:, :, l]))
        return r["agent_steps"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, n_threads)) as pool:
        steps = sum(pool.map(one, list(leaf_ids)))
    return steps, time.perf_counter() - t0


def cpu_sample(kind, w, target_seconds):
    """How many leaves (with all their rollouts) the CPU arm runs per step: about `target_seconds`
    of work, judged from a probe of two leaves."""
    cores = os.cpu_count() or 1
    n = w["rollouts_per_leaf"]
    probe = list(range(min(cores, w["n_leaves"])))
    steps, sec = cpu_rollouts(kind, w, probe, n, cores, 1000, 0)
    per_leaf = max(sec / len(probe), 1e-6)          # wall time per leaf with all cores busy
    k = int(max(cores, min(w["n_leaves"], target_seconds / per_leaf)))
    return min(w["n_leaves"], (k + cores - 1) // cores * cores)


def cpu_baseline(kind, rank, target_seconds=12.0):
    from planner_rules_mcts_b200 import parallel
    w = reference_workload(kind, rank)
    cores = os.cpu_count() or 1
    k = cpu_sample(kind, w, target_seconds)
    base = parallel.rollout_base(rank, 0, w["n_leaves"] * w["rollouts_per_leaf"])
    steps, sec = cpu_rollouts(kind, w, range(k), w["rollouts_per_leaf"], cores, 1000, base)
    return dict(value=steps / sec, unit=UNIT, cores=cores, kind=w["kind_name"],
                sample="leaves 0..%d of the decision's %d (same leaves, same rollout ids as the GPU arm), %d rollouts "
                       "each, on %d host threads (%.1f s)" % (k - 1, w["n_leaves"], w["rollouts_per_leaf"], cores, sec))


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The CPU implementation of the path on the host cores: oracle/_ref (the reference's own
    GridWorld sources) for workload G, the FP64 restatement in oracle/ for S / M / W.  Only the
    oracle is built and loaded here -- never libbrm.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/gridworld"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    from planner_rules_mcts_b200 import parallel
    kind = args.workload
    w = reference_workload(kind, 0)
    cores = os.cpu_count() or 1
    k = cpu_sample(kind, w, 4.0)        # leaves per step: a bounded sample of the decision
    n = w["rollouts_per_leaf"]
    per_decision = w["n_leaves"] * n
    for i in range(args.warmup):
        cpu_rollouts(kind, w, range(min(2, k)), n, cores, 1000, parallel.rollout_base(0, i, per_decision))
    total_steps, total_sec = 0, 0.0
    for i in range(args.steps):
        s, sec = cpu_rollouts(kind, w, range(k), n, cores, 1000, parallel.rollout_base(0, args.warmup + i, per_decision))
        total_steps += s
        total_sec += sec
    value = total_steps / max(total_sec, 1e-9)
    sample = ("leaves 0..%d of the decision's %d, %d rollouts each per step (same leaves and rollout ids as the GPU "
              "arm), x %d steps on %d host threads" % (k - 1, w["n_leaves"], n, args.steps, cores))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total_sec / max(args.steps, 1), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64" if kind != "G" else "i32", data="synthetic", impl="reference",
                config=dict(workload=w["name"], description=w["desc"], leaves=w["n_leaves"], rollouts_per_leaf=n,
                            max_steps=w["max_steps"], same_leaves_as_gpu_arm=True,
                            note="each reference step is a bounded sample of the decision: " + sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=w["kind_name"], sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def kernel_source_hash():
    """sha256 over the CUDA sources the library is built from: ties profiles/inst_counts.json (ncu
    instruction counts per agent-step) to the kernels that are running."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "planner-rules-mcts_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def load_inst_counts(kind):
    path = os.path.join(ROOT, "profiles", "inst_counts.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            stale = d.get("csrc_sha16") != kernel_source_hash()
            for v in d.values():
                if isinstance(v, dict):
                    v["stale"] = stale
            if kind != "R":
                return d.get(kind)
            # R runs M's kernel on M's scenario: per-agent-step counts carry over, per-launch traffic does not
            m = dict(d.get("M") or {})
            if m:
                m["dram_bytes_per_launch"] = None
                m["source"] = "workload M's capture (same kernel and scenario): " + m.get("source", "")
            return m or None
        except (ValueError, OSError):
            return None
    return None


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rollout engine has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if world > 1:
        # NCCL_DEBUG=VERSION prints a banner on stdout, which must carry only the JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import __graft_entry__ as g
    g.build(only_missing=True)
    from planner_rules_mcts_b200 import parallel

    kind = args.workload
    # a dedicated (non-default) torch stream: the engine enqueues on it, and the torch events
    # below are recorded on it, so they bracket exactly the engine's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sampler = ClockSampler(local_rank)
    sampler.start()
    w = build_workload(kind, local_rank, rank, cuda_stream=stream.cuda_stream)
    eng, L, n = w["engine"], w["leaves"], w["rollouts_per_leaf"]
    # root-parallel merge behind the C ABI: with more than one rank every brm_*_rollout_root_stats
    # call returns statistics summed over the box (NCCL on the engine's stream, before the D2H)
    parallel.init_comm(eng.engine)
    n_leaves = L.n
    M, V, R, A = w["M"], w["V"], w["R"], w["A"]
    rollouts_per_step = n_leaves * n

    # device-resident inputs and outputs
    d_leaves = L.cuda(local_rank)
    d_first = torch.from_numpy(w["first"]).to(dev)
    out = dict(returns_sum=torch.zeros((n_leaves, M, V), dtype=torch.float64, device=dev),
               viol_sum=torch.zeros((n_leaves, M, R), dtype=torch.int64, device=dev),
               agent_steps=torch.zeros(n_leaves, dtype=torch.int64, device=dev))
    stats = torch.zeros((M, w["n_actions"], 1 + V), dtype=torch.float64, device=dev)
    steps_acc = torch.zeros((), dtype=torch.int64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def rollout_kwargs(step_idx):
        base = parallel.rollout_base(rank, step_idx, rollouts_per_step)  # disjoint Philox streams
        kw = dict(seed=1000, rollout_base=base, max_steps=w["max_steps"], discount_factor=w["gamma"])
        return kw

    def device_step(step_idx):
        stats.zero_()
        # one C-ABI call: rollouts, per-leaf sums, root statistics and (world > 1) their merge
        eng.rollout(d_leaves, n, out=out, first_action=d_first, stats=stats, **rollout_kwargs(step_idx))
        steps_acc.add_(out["agent_steps"].sum())

    for i in range(args.warmup):
        device_step(i)
    torch.cuda.synchronize(dev)
    steps_acc.zero_()
    eng.engine.set_kernel_timing(True)
    eng.engine.kernel_time(reset=True)
    launches0 = eng.engine.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t_wall0 = time.perf_counter()
    evs = []
    for i in range(args.steps):
        flush.fill_(i & 0xff)  # L2 flush between timed iterations (outside the event pair)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        device_step(args.warmup + i)
        b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    clocks = sampler.stop(t_wall0, t_wall1)
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    kern_ms, kern_launches = eng.engine.kernel_time(reset=True)
    eng.engine.set_kernel_timing(False)
    gpu_launches = eng.engine.launch_count - launches0
    agent_steps = int(steps_acc.item())

    # ---- end-to-end: public API with HOST buffers (pinned), H2D + D2H inside the timed region
    def pin(a):
        # pinned host copy (torch has no uint32/uint64 arithmetic: pin the signed view)
        a = np.ascontiguousarray(a)
        signed = {np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32}.get(a.dtype)
        t = torch.from_numpy(a.view(signed) if signed else a).pin_memory()
        return t.numpy().view(a.dtype) if signed else t.numpy()
    fields = type(L).FIELDS if hasattr(type(L), "FIELDS") else ("agents", "term", "depth", "q", "viol")
    host_leaves = type(L)(*[None if getattr(L, f) is None else pin(getattr(L, f)) for f in fields])
    h_first = pin(w["first"])
    h_out = dict(returns_sum=pin(np.zeros((n_leaves, M, V), np.float64)),
                 viol_sum=pin(np.zeros((n_leaves, M, R), np.uint64)),
                 agent_steps=pin(np.zeros(n_leaves, np.uint64)))
    h_stats = pin(np.zeros((M, w["n_actions"], 1 + V), np.float64))
    h2d = sum(getattr(host_leaves, f).nbytes for f in fields if getattr(host_leaves, f) is not None)
    h2d += h_first.nbytes + h_stats.nbytes  # first actions and the statistics accumulated into
    d2h = sum(v.nbytes for v in h_out.values()) + h_stats.nbytes

    def e2e_step(step_idx):
        h_stats[:] = 0
        # one C-ABI call per decision: rollouts + root statistics, one upload / download / synchronisation
        # (with more than one rank the statistics come back merged: the allreduce runs on the
        # device inside the call, between the root-statistics kernel and the download)
        eng.rollout(host_leaves, n, out=h_out, first_action=h_first, stats=h_stats, **rollout_kwargs(step_idx))
        return int(h_out["agent_steps"].sum())

    e2e_step(10_000)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_steps_n = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    e2e_agent_steps = 0
    for i in range(e2e_steps_n):
        e2e_agent_steps += e2e_step(20_000 + i)
    torch.cuda.synchronize(dev)
    e2e_sec = time.perf_counter() - t0

    # ---- aggregate over ranks: totals summed, times maxed
    if world > 1:
        t = torch.tensor([dev_ms, e2e_sec, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, e2e_sec_max, kern_ms_max = [float(x) for x in t.tolist()]
        c = torch.tensor([agent_steps, e2e_agent_steps, gpu_launches], dtype=torch.int64, device=dev)
        dist.all_reduce(c)
        agent_steps_all, e2e_steps_all, launches_all = [int(x) for x in c.tolist()]
    else:
        dev_ms_max, e2e_sec_max, kern_ms_max = dev_ms, e2e_sec, kern_ms
        agent_steps_all, e2e_steps_all, launches_all = agent_steps, e2e_agent_steps, gpu_launches

    if rank == 0:
        value = agent_steps_all / (dev_ms_max * 1e-3)
        info = eng.engine.device_info()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # dominant kernel = the fused rollout kernel (+ its tiny leaf reduce), timed live with
        # CUDA events on the launching stream (brm_kernel_time)
        kern_s = kern_ms_max * 1e-3
        per_rank_steps = agent_steps_all / world
        # algorithmic bytes of one rollout launch (rollout mode): every rollout
        # reads its leaf state + rule states + Philox key/counter and writes returns and
        # violation counters; lane polylines / automaton tables are staged once per CTA
        sb = w["state_bytes"]
        bytes_per_rollout = A * (sb + 1 + R) + 16 + M * ((sb // 4) * V + 4 * R)
        algo_bytes = rollouts_per_step * bytes_per_rollout * args.steps
        hbm_achieved = algo_bytes / kern_s / 1e9
        sm_mhz = clocks.get("sm_mhz") or info["sm_clock_khz"] / 1e3
        issue_peak = info["sm_count"] * 4 * sm_mhz * 1e6  # warp-instructions / s at the sampled clock
        ic = load_inst_counts(kind)
        roof = dict(bound="issue", unit="Gwarp-inst/s", peak=issue_peak / 1e9,
                    peak_note="%d SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the run)"
                              % (info["sm_count"], sm_mhz),
                    kernel="cw_pool_kernel<ROLLOUT>" if kind != "G" else "gw_rollout_kernel",
                    kernel_ms_per_launch=kern_ms_max / max(args.steps, 1),
                    kernel_share_of_step=kern_ms_max / dev_ms_max if dev_ms_max else None,
                    hbm=dict(bound="hbm", achieved=hbm_achieved, peak=hbm_peak, unit="GB/s",
                             frac=hbm_achieved / hbm_peak, peak_source=hbm_src,
                             algorithmic_bytes_per_launch=rollouts_per_step * bytes_per_rollout,
                             traffic=(ic or {}).get("dram_bytes_per_launch")))
        if ic and ic.get("warp_inst_per_agent_step"):
            achieved = per_rank_steps * ic["warp_inst_per_agent_step"] / kern_s
            # the issue roofline on THREAD instructions (peak x 32 lanes); the gap to
            # `frac` is the share of lanes that are masked off when an instruction issues
            thread_rate = per_rank_steps * ic["thread_inst_per_agent_step"] / kern_s
            roof.update(achieved=achieved / 1e9, frac=achieved / issue_peak,
                        frac_thread=thread_rate / (issue_peak * 32.0),
                        avg_active_lanes=ic.get("avg_active_lanes"),
                        warp_inst_per_agent_step=ic["warp_inst_per_agent_step"],
                        thread_inst_per_agent_step=ic.get("thread_inst_per_agent_step"),
                        inst_source=ic.get("source"), stale=bool(ic.get("stale")),
                        traffic=ic.get("dram_bytes_per_launch"))
        else:
            roof.update(achieved=None, frac=None, traffic=None,
                        inst_source="no ncu instruction count committed yet for this workload")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=dev_ms_max / max(args.steps, 1), higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype=w["dtype"], data="synthetic",
                    config=dict(workload=w["name"], description=w["desc"], leaves_per_gpu=n_leaves,
                                rollouts_per_leaf=n, rollouts_per_decision_per_gpu=rollouts_per_step,
                                max_steps=w["max_steps"], agents=A, mcts_agents=M, reward_vec_size=V,
                                rule_states_per_agent=R, parallelism="root-parallel x%d" % world,
                                l2="flushed between timed iterations (256 MiB fill)",
                                rollouts_per_sec=rollouts_per_step * world * args.steps / (dev_ms_max * 1e-3),
                                decisions_per_sec=world * args.steps / (dev_ms_max * 1e-3),
                                agent_steps_per_decision_per_gpu=per_rank_steps / max(args.steps, 1)),
                    clocks=clocks,
                    e2e=dict(value=e2e_steps_all / e2e_sec_max, unit=UNIT, h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=int(d2h), steps=e2e_steps_n,
                             note="planner_rules_mcts_b200 engine.rollout with root statistics (one "
                                  "brm_*_rollout_root_stats call per decision) on pinned host buffers; H2D, "
                                  "kernels, D2H and the host wall clock all inside"),
                    gpu_launches=int(launches_all), roofline=roof, wall_s=t_wall)
        line["comm"] = dict(ranks=eng.engine.comm_size(), where="brm_comm_init / NCCL inside brm_*_rollout_root_stats")
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(kind, rank)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="M", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
