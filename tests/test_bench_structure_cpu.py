"""bench.py under torchrun: sections that only rank 0 runs must not issue a collective. The engine all-reduces the
root statistics / rollout moments whenever a communicator is attached, so a rank-0-only section may only run after
engine.comm_destroy() - a run with 8 ranks once hung on exactly this (a moments request in a rank-0-only call).
Checked on the source, since the multi-rank run itself needs GPUs."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _source():
    with open(os.path.join(ROOT, "bench.py")) as f:
        return f.read()


def _function_source(src, name):
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise AssertionError(f"bench.py has no function {name}")


def test_rank0_only_sections_run_after_the_communicator_is_gone():
    src = _source()
    main = _function_source(src, "run_native") if "def run_native" in src else src
    destroy = main.index("engine.comm_destroy()")
    for call in ("_c3_variants(engine", "measure_c4_hypothesis(engine"):
        at = main.index(call)
        assert at > destroy, f"{call} must come after engine.comm_destroy()"
    # and the sharded sections (every rank calls them) come before it
    for call in ("measure_c5_strong(engine", "measure_rollout_c3(engine"):
        assert main.index(call) < destroy, call


def test_c3_variants_never_ask_for_the_all_reduced_moments():
    body = _function_source(_source(), "_c3_variants")
    code = "\n".join(l for l in body.splitlines() if not l.strip().startswith(("#", '"""')))
    assert "d_moments=" not in code and "moments=True" not in code
    # nor does the sharded C3 section call it (it used to, on rank 0, with the communicator attached)
    assert "_c3_variants(" not in _function_source(_source(), "measure_rollout_c3")
