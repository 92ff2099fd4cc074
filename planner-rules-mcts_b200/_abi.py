"""ctypes view of the C ABI declared in include/brm.h (libbrm.so).

The library is the product; this module only loads it and declares the argument
types.  There is no CPU fallback: if libbrm.so is missing the import fails, and every
compute entry point fails with BRM_ERR_NO_DEVICE when no B200-class GPU is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BRM_LIB_PATH selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("BRM_LIB_PATH") or os.path.join(HERE, "libbrm.so")

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_TERMINAL, ERR_NOT_CONFIGURED = 0, -1, -2, -3, -4, -5
MEM_HOST, MEM_DEVICE = 0, 1
MAX_APS, MAX_DFA_STATES, MAX_RULES, MAX_REWARD, MAX_ACTIONS = 4, 16, 8, 4, 8
VIOLATION = 0xFF
GW_MAX_AGENTS, GW_MAX_SLOTS = 8, 16
GW_LABEL_COLLISION, GW_LABEL_MERGED_E, GW_LABEL_MERGED_X, GW_LABEL_IN_DIRECT_FRONT_X = 0, 1, 2, 3


class BrmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("brm error %d: %s" % (code, msg))
        self.code = code


class Rule(C.Structure):
    _fields_ = [
        ("n_aps", C.c_uint8), ("n_states", C.c_uint8), ("init_state", C.c_uint8),
        ("priority", C.c_uint8), ("on_violation", C.c_uint8), ("_pad", C.c_uint8 * 3),
        ("weight", C.c_float),
        ("ap_label", C.c_uint8 * MAX_APS),
        ("ap_agent_specific", C.c_uint8 * MAX_APS),
        ("trans", C.c_uint8 * (MAX_DFA_STATES * (1 << MAX_APS))),
        ("accepting", C.c_uint8 * MAX_DFA_STATES),
    ]


class RolloutParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("rollout_base", C.c_uint64),
        ("rollouts_per_leaf", C.c_int32), ("max_steps", C.c_int32),
        ("discount_factor", C.c_double),
        ("first_discount_exp", C.c_int32), ("_pad", C.c_int32),
        ("coop_factor", C.c_double),
    ]


class GwParams(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32), ("state_x_length", C.c_int32),
        ("ego_goal_reached_position", C.c_int32), ("terminal_depth", C.c_int32),
        ("merging_point", C.c_int32), ("reward_vec_size", C.c_int32),
        ("n_actions", C.c_int32), ("action_map", C.c_int32 * MAX_ACTIONS),
        ("speed_deviation_weight", C.c_double), ("acceleration_weight", C.c_double),
        ("potential_weight", C.c_double),
    ]


class GwStates(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("mem", C.c_int32),
        ("agents", C.c_void_p), ("term", C.c_void_p), ("depth", C.c_void_p),
        ("q", C.c_void_p), ("viol", C.c_void_p),
    ]


class GwRolloutOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "returns_sum", "viol_sum", "agent_steps", "ro_returns", "ro_viol", "ro_q", "ro_agents",
        "ro_steps")]


class GwTrace(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "agents_out", "term_out", "q_out", "viol_out", "rewards_out", "labels_out",
        "term_reward_out", "steps_out")]


CW_MAX_AGENTS, CW_MAX_LANES, CW_MAX_PTS, CW_MAX_ROAD, CW_MAX_PRIMS, CW_MAX_SLOTS = 16, 8, 64, 32, 8, 32
(CW_LABEL_COLLISION_EGO, CW_LABEL_COLLISION_CORRIDOR, CW_LABEL_GOAL_REACHED,
 CW_LABEL_SAFE_DISTANCE_FRONT, CW_LABEL_MERGED_E, CW_LABEL_RIGHTMOST_LANE, CW_LABEL_ON_ROAD) = range(7)
CW_LABEL_MERGED_X, CW_LABEL_IN_DIRECT_FRONT_X, CW_LABEL_NEAR_X = 8, 9, 10
CW_PRESENT, CW_VALID = 1, 2


class CwParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "collision_weight", "out_of_map_weight", "potential_weight", "acceleration_weight",
        "radial_acceleration_weight", "desired_velocity_weight", "lane_center_weight",
        "prediction_time_span", "desired_velocity", "discount_factor", "goal_reward")] + [
        ("reward_vector_size", C.c_int32), ("horizon", C.c_int32),
        ("use_rule_reward_for_ego_only", C.c_int32), ("remove_agents_out_of_map", C.c_int32)] + [
        (k, C.c_double) for k in (
            "wheel_base", "integration_time_delta", "sd_reaction_time", "sd_max_decel_rear",
            "sd_max_decel_front", "near_distance")]


class CwAgent(C.Structure):
    _fields_ = [("front", C.c_double), ("rear", C.c_double), ("half_width", C.c_double),
                ("goal", C.c_double * 6), ("has_goal", C.c_int32), ("n_prims", C.c_int32),
                ("prim_acc", C.c_double * CW_MAX_PRIMS), ("prim_delta", C.c_double * CW_MAX_PRIMS)]


class CwLane(C.Structure):
    _fields_ = [("n_pts", C.c_int32), ("successor", C.c_int32), ("flags", C.c_int32), ("_pad", C.c_int32),
                ("half_width", C.c_double), ("s_point", C.c_double),
                ("pts", (C.c_double * 2) * CW_MAX_PTS)]


class CwScenario(C.Structure):
    _fields_ = [("params", CwParams), ("n_agents", C.c_int32), ("n_mcts_agents", C.c_int32),
                ("n_lanes", C.c_int32), ("n_road", C.c_int32), ("n_rules", C.c_int32), ("_pad", C.c_int32),
                ("agents", CwAgent * CW_MAX_AGENTS), ("lanes", CwLane * CW_MAX_LANES),
                ("road", (C.c_double * 2) * CW_MAX_ROAD), ("rules", Rule * MAX_RULES)]


class CwStates(C.Structure):
    _fields_ = [("n", C.c_int32), ("mem", C.c_int32)] + [(k, C.c_void_p) for k in (
        "scenario", "agents", "flags", "horizon", "terminal", "q", "viol")]


class CwRolloutOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "returns_sum", "viol_sum", "agent_steps", "ro_returns", "ro_viol", "ro_q", "ro_agents",
        "ro_flags", "ro_steps")]


class CwTrace(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "agents_out", "flags_out", "terminal_out", "q_out", "viol_out", "rewards_out", "labels_out",
        "frenet_out", "term_reward_out", "steps_out")]


# every symbol include/brm.h declares (tests/test_abi.py checks the library exports them)
SYMBOLS = [
    "brm_cw_configure", "brm_cw_rules_per_agent", "brm_cw_check_terminal", "brm_cw_execute",
    "brm_cw_evaluate_rules", "brm_cw_rollout_root_stats", "brm_gw_rollout_root_stats",
    "brm_cw_get_num_actions", "brm_cw_terminal_reward", "brm_cw_rollout", "brm_cw_replay",
    "brm_create", "brm_destroy", "brm_last_error", "brm_abi_version", "brm_abi_struct_sizes", "brm_synchronize",
    "brm_launch_count", "brm_set_kernel_timing", "brm_kernel_time", "brm_device_info",
    "brm_gw_configure", "brm_gw_rules_per_agent", "brm_gw_execute", "brm_gw_get_num_actions",
    "brm_gw_is_terminal", "brm_gw_terminal_reward", "brm_gw_rollout", "brm_gw_replay",
    "brm_root_stats_reduce", "brm_comm_unique_id", "brm_comm_init", "brm_comm_destroy", "brm_comm_size",
    "brm_allreduce_root_stats",
]

_lib = None


def load():
    """Load libbrm.so (built in-tree by __graft_entry__.build()); fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()'); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64, dbl = C.c_void_p, C.c_int32, C.c_uint64, C.c_double
    P = C.POINTER
    lib.brm_create.argtypes = [C.c_int, vp, P(vp)]
    lib.brm_destroy.argtypes = [vp]
    lib.brm_destroy.restype = None
    lib.brm_last_error.restype = C.c_char_p
    lib.brm_synchronize.argtypes = [vp]
    lib.brm_launch_count.argtypes = [vp]
    lib.brm_launch_count.restype = u64
    lib.brm_set_kernel_timing.argtypes = [vp, C.c_int]
    lib.brm_kernel_time.argtypes = [vp, P(dbl), P(u64), C.c_int]
    lib.brm_device_info.argtypes = [vp, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int)]
    lib.brm_gw_configure.argtypes = [vp, P(GwParams), P(Rule), i32]
    lib.brm_gw_rules_per_agent.argtypes = [vp]
    lib.brm_gw_execute.argtypes = [vp, P(GwStates), vp, P(GwStates), vp]
    lib.brm_gw_get_num_actions.argtypes = [vp, P(GwStates), vp]
    lib.brm_gw_is_terminal.argtypes = [vp, P(GwStates), vp]
    lib.brm_gw_terminal_reward.argtypes = [vp, P(GwStates), vp]
    lib.brm_gw_rollout.argtypes = [vp, P(GwStates), P(RolloutParams), P(GwRolloutOut)]
    lib.brm_gw_replay.argtypes = [vp, P(GwStates), vp, i32, P(GwTrace)]
    lib.brm_root_stats_reduce.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp]
    lib.brm_cw_configure.argtypes = [vp, P(CwScenario), i32]
    lib.brm_cw_rules_per_agent.argtypes = [vp]
    lib.brm_cw_check_terminal.argtypes = [vp, P(CwStates)]
    lib.brm_cw_execute.argtypes = [vp, P(CwStates), vp, P(CwStates), vp]
    lib.brm_cw_evaluate_rules.argtypes = [vp, P(CwStates), P(CwStates), vp]
    lib.brm_cw_get_num_actions.argtypes = [vp, P(CwStates), vp]
    lib.brm_cw_terminal_reward.argtypes = [vp, P(CwStates), vp]
    lib.brm_cw_rollout.argtypes = [vp, P(CwStates), P(RolloutParams), P(CwRolloutOut)]
    lib.brm_cw_rollout_root_stats.argtypes = [vp, P(CwStates), P(RolloutParams), P(CwRolloutOut), vp, i32, vp]
    lib.brm_gw_rollout_root_stats.argtypes = [vp, P(GwStates), P(RolloutParams), P(GwRolloutOut), vp, i32, vp]
    lib.brm_cw_replay.argtypes = [vp, P(CwStates), vp, i32, P(CwTrace)]
    lib.brm_comm_unique_id.argtypes = [vp]
    lib.brm_comm_init.argtypes = [vp, vp, i32, i32]
    lib.brm_comm_destroy.argtypes = [vp]
    lib.brm_comm_size.argtypes = [vp]
    lib.brm_allreduce_root_stats.argtypes = [vp, i32, vp, C.c_int64]
    _lib = lib
    return lib


def _prefer_torch_nccl():
    """libbrm.so binds NCCL at run time (dlopen "libnccl.so.2").  In a Python process that also
    uses PyTorch there must be ONE NCCL, and it has to be the one PyTorch was built against (its
    wheel bundles it): point BRM_NCCL_LIB at that copy unless the caller chose one."""
    if os.environ.get("BRM_NCCL_LIB"):
        return
    import importlib.util
    for pkg in ("nvidia.nccl", "torch"):
        try:
            spec = importlib.util.find_spec(pkg)
        except (ImportError, ValueError):
            spec = None
        if not spec or not spec.submodule_search_locations:
            continue
        base = list(spec.submodule_search_locations)[0]
        for cand in (os.path.join(base, "lib", "libnccl.so.2"),
                     os.path.join(base, "..", "nvidia", "nccl", "lib", "libnccl.so.2")):
            if os.path.exists(cand):
                os.environ["BRM_NCCL_LIB"] = os.path.abspath(cand)
                return


def comm_unique_id():
    """128 bytes identifying a new communicator (brm_comm_unique_id; call on one rank)."""
    _prefer_torch_nccl()
    buf = (C.c_uint8 * 128)()
    check(load().brm_comm_unique_id(buf))
    return bytes(buf)


def check(rc):
    if rc != OK:
        raise BrmError(rc, load().brm_last_error().decode("utf-8", "replace"))


class Engine:
    """One GPU + one CUDA stream (brm_engine)."""

    def __init__(self, device=0, cuda_stream=None):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.brm_create(int(device), C.c_void_p(cuda_stream or 0), C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.brm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self.lib.brm_synchronize(self.handle))

    # -- root-parallel communicator (NCCL behind the ABI) ------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        """brm_comm_init: collective over the ranks that share `unique_id` (made on rank 0 with
        `comm_unique_id()` and handed round by the caller)."""
        _prefer_torch_nccl()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self.lib.brm_comm_init(self.handle, buf, int(rank), int(n_ranks)))

    def comm_size(self):
        return int(self.lib.brm_comm_size(self.handle))

    def comm_destroy(self):
        check(self.lib.brm_comm_destroy(self.handle))

    def allreduce_root_stats(self, stats):
        """In-place sum over the ranks of a [..] float64 buffer (numpy or torch CUDA tensor)."""
        dev = type(stats).__module__.startswith("torch")
        ptr = stats.data_ptr() if dev else stats.ctypes.data
        n = stats.numel() if dev else stats.size
        check(self.lib.brm_allreduce_root_stats(self.handle, MEM_DEVICE if dev else MEM_HOST, ptr, int(n)))
        return stats

    @property
    def launch_count(self):
        return int(self.lib.brm_launch_count(self.handle))

    def set_kernel_timing(self, enabled=True):
        check(self.lib.brm_set_kernel_timing(self.handle, 1 if enabled else 0))

    def kernel_time(self, reset=True):
        ms, n = C.c_double(0), C.c_uint64(0)
        check(self.lib.brm_kernel_time(self.handle, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, int(n.value)

    def device_info(self):
        a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.lib.brm_device_info(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(sm_count=a.value, sm_clock_khz=b.value, cc=(c.value, d.value))
