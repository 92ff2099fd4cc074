"""planner_rules_mcts_b200 -- B200-native rollout engine for rule-aware multi-agent MCTS.

Drop-in for the simulation phase of bark-simulator/planner-rules-mcts: the state
plugins (GridWorldState, MvmctsState) and the RandomHeuristic rollout loop run as
hand-written sm_100a CUDA kernels behind the C ABI of include/brm.h (libbrm.so, built
in-tree).  This package is the thin host layer mirroring the reference's interface.
The directory is named `planner-rules-mcts_b200`; import it as
`planner_rules_mcts_b200` (loader module at the repository root).
"""
from . import _abi, ltl  # noqa: F401
from ._abi import BrmError, Engine  # noqa: F401
from .gridworld import (GridWorldEngine, GridWorldState, GridWorldStateParameter,  # noqa: F401
                        GwBatch, ZIP_FORMULA, make_default_rules)
from .ltl import RuleMonitor, compile_rule  # noqa: F401

__all__ = ["Engine", "BrmError", "GridWorldEngine", "GridWorldState", "GridWorldStateParameter",
           "GwBatch", "RuleMonitor", "compile_rule", "ZIP_FORMULA", "make_default_rules", "ltl"]
from .continuous import (AgentModel, CwBatch, Lane, MvmctsState, RulesMctsEngine,  # noqa: E402,F401
                         RulesMctsStateParameters, Scenario)
from . import behavior, mcts, parallel, planner, scenarios  # noqa: E402,F401

__all__ += ["AgentModel", "CwBatch", "Lane", "MvmctsState", "RulesMctsEngine", "RulesMctsStateParameters",
            "Scenario", "scenarios", "planner", "parallel", "mcts", "behavior"]
