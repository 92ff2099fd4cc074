"""The C-ABI library loads and exports every symbol include/brm.h declares; without a GPU
the engine refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "brm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(brm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(_abi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libbrm.so does not export %s" % s
    assert sorted(_abi.SYMBOLS) == syms, "planner_rules_mcts_b200._abi.SYMBOLS out of sync with brm.h"
    assert lib.brm_abi_version() == 3


def test_struct_sizes_match_header(built):
    # sizes the C side was compiled with (computed from the header's layout rules)
    assert ctypes.sizeof(_abi.Rule) == 8 + 4 + 4 + 4 + 256 + 16
    assert ctypes.sizeof(_abi.RolloutParams) == 48
    assert ctypes.sizeof(_abi.GwParams) == 7 * 4 + 8 * 4 + 4 + 3 * 8
    assert ctypes.sizeof(_abi.GwStates) == 8 + 5 * 8


def test_ctypes_layouts_match_the_compiled_library(built):
    """every struct of the ctypes binding has the size the library was compiled with"""
    lib = ctypes.CDLL(_abi.LIB_PATH)
    out = (ctypes.c_int32 * 16)()
    n = lib.brm_abi_struct_sizes(out, 16)
    mine = [_abi.Rule, _abi.RolloutParams, _abi.GwParams, _abi.GwStates, _abi.CwParams, _abi.CwAgent, _abi.CwLane,
            _abi.CwScenario, _abi.CwStates, _abi.CwRolloutOut, _abi.CwTrace]
    assert n == len(mine)
    assert [ctypes.sizeof(t) for t in mine] == list(out[:n])


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(brm.BrmError) as ei:
        brm.Engine(0)
    assert ei.value.code == _abi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)
