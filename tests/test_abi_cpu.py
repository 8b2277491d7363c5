"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the public
header declares, its host-only entry points agree with the reference defaults, and compute entry
points fail loudly (no CPU fallback) when there is no CUDA device."""
from __future__ import annotations

import ctypes as C
import os
import re

import pytest

from helpers import golden
from mamcts_b200 import _abi, default_crossing_params, default_params, default_simple_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_abi.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return _abi.load()


def test_library_exports_every_declared_symbol(lib):
    with open(os.path.join(ROOT, "include", "mamcts_b200.h")) as f:
        header = f.read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mamcts_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libmamcts_b200.so does not export {name}"
    assert sorted(_abi.EXPORTED_SYMBOLS) == declared
    assert lib.mamcts_abi_version() == 1


def test_struct_layouts_match_header_sizes(lib):
    # sizes computed by hand from include/mamcts_b200.h (LP64): a ctypes / C mismatch would corrupt
    # every call, so pin them.
    assert C.sizeof(_abi.Params) == 6 * 8 + 8 * 4
    assert C.sizeof(_abi.CrossingParams) == 3 * 8 + 10 * 4
    assert C.sizeof(_abi.SimpleParams) == 8
    assert C.sizeof(_abi.ActionSource) == 8 + 24 + 8 + 16 * 4 + 8 + 8 + 8 + 8 + 16 + 16
    assert C.sizeof(_abi.RolloutOut) == 8 + 12 * 8 + 8 + 8
    assert C.sizeof(_abi.UctStats) == 16 + 5 * 8


def test_defaults_match_reference_defaults(lib):
    p = _abi.Params()
    lib.mamcts_default_params(C.byref(p))
    q = default_params()
    for name, _ in _abi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    # reference mcts/mcts_parameters.h:97-116
    assert (p.discount_factor, p.random_seed, p.max_number_of_iterations) == (0.9, 1000, 10000)
    assert (p.uct_lower_bound, p.uct_upper_bound, p.uct_exploration_constant) == (-1000.0, 100.0, 0.7)
    assert (p.uct_pw_k, p.uct_pw_alpha, p.rh_max_number_of_iterations) == (4.0, 0.25, 1000)
    c = _abi.CrossingParams()
    lib.mamcts_default_crossing_params(C.byref(c))
    d = default_crossing_params()
    for name, _ in _abi.CrossingParams._fields_:
        assert getattr(c, name) == getattr(d, name), name
    # reference environments/crossing_state_parameters.h:36-59
    assert (c.num_other_agents, c.chain_length, c.ego_goal_pos, c.num_other_actions) == (2, 21, 12, 7)
    assert d.crossing_point == 11 and d.num_ego_actions == 4
    s = _abi.SimpleParams()
    lib.mamcts_default_simple_params(C.byref(s))
    assert (s.winning_state_length, s.loosing_state_length) == (10, -1)
    assert default_simple_params().winning_state_length == 10


def test_expand_order_is_libstdcxx_order(lib):
    # mamcts_reference_expand_order is host-only: the reference's deterministic widening order
    for seed, per_n in golden("expand_order.json").items():
        for n, order in per_n.items():
            out = (C.c_uint32 * int(n))()
            assert lib.mamcts_reference_expand_order(int(seed), int(n), out) == 0
            assert list(out) == order
    out = (C.c_uint32 * 4)()
    assert lib.mamcts_reference_expand_order(1000, 0, out) != 0
    assert b"num_actions" in lib.mamcts_last_error(None)


def _have_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_ctypes_structs_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc sees include/mamcts_b200.h, against the ctypes mirrors."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    structs = {"mamcts_config": _abi.Config, "mamcts_params": _abi.Params,
               "mamcts_crossing_params": _abi.CrossingParams, "mamcts_simple_params": _abi.SimpleParams,
               "mamcts_crossing_params_f32": _abi.CrossingParamsF32,
               "mamcts_action_source": _abi.ActionSource, "mamcts_rollout_out": _abi.RolloutOut,
               "mamcts_uct_stats": _abi.UctStats, "mamcts_hyp_stats": _abi.HypStats,
               "mamcts_search_config": _abi.SearchConfig}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mamcts_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _t in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi_sizes.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _t in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


@pytest.mark.skipif(_have_cuda(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu(lib):
    h = C.c_void_p()
    cfg = _abi.Config(device=0, flags=0)
    rc = lib.mamcts_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.mamcts_last_error(None)
    from mamcts_b200.engine import Engine, MamctsError
    with pytest.raises(MamctsError):
        Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mamcts_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "mamcts_oracle" not in text and "oracle_api" not in text, fn
                assert "libmamcts_ref" not in text, fn
