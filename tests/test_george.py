"""Parity against the REFERENCE's own solver, george 0.2.x (gprot/model.py:9-10, 334, 343, 346, 354).

george is not installable in the build container, so these tests are opportunistic:

* with a live george 0.2.x (pytest.importorskip): the oracle -- and on the GPU box the CUDA path -- against
  george.GP(k, solver=BasicSolver) and solver=HODLRSolver, the reference's call sequence
  (tests/golden/make_george_golden.py::george_lnlike);
* with a committed tests/golden/george_golden.npz (written by that script wherever george exists): the same
  comparison against the stored reference-held values, no george needed.

Bounds: BasicSolver is the same dense factorisation as the oracle -> 1e-9 relative (north star) with the condition
number of the case scaling it above 1e8; HODLRSolver approximates the matrix to tol = 1e-12 -> tol * cond.
Until one of the two sources exists every test here skips and the repository says PARITY UNPINNED.
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import gp_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "george_golden.npz")
CASES = ["n24", "n64", "n150", "n300c100", "n257"]


def _maker():
    spec = importlib.util.spec_from_file_location("make_george_golden", os.path.join(HERE, "golden", "make_george_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _bound(solver, cond):
    cond = np.asarray(cond, dtype=np.float64)
    if solver == "BasicSolver":
        return 1e-9 * np.maximum(1.0, cond / 1e8)
    return np.maximum(1e-9, 1e-12 * cond)            # HODLR: tol * cond(K)


def _reference_values(name, solver, golden):
    """(values, source) from a live george if importable, else from the committed file, else skip."""
    try:
        import george  # noqa: F401

        mk = _maker()
        lists = mk.chunk_lists(golden[name + "_t"], golden[name + "_y"], golden[name + "_yerr"], int(golden[name + "_chunksize"]))
        return np.array([mk.george_lnlike(th, *lists, solver_name=solver) for th in golden[name + "_theta"]]), "live george"
    except ImportError:
        pass
    if os.path.exists(GOLD):
        return np.load(GOLD)["%s_%s" % (name, solver)], "tests/golden/george_golden.npz"
    pytest.skip("no george 0.2.x and no tests/golden/george_golden.npz: PARITY UNPINNED at the george boundary "
                "(run tests/golden/make_george_golden.py where george<0.3 is installed)")


def _compare(got, ref, bound, what):
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.array_equal(np.isfinite(got), np.isfinite(ref)), what
    m = np.isfinite(ref)
    rel = np.abs(got[m] - ref[m]) / np.abs(ref[m])
    assert np.all(rel <= np.asarray(bound)[m]), (what, rel, np.asarray(bound)[m])


@pytest.mark.parametrize("solver", ["BasicSolver", "HODLRSolver"])
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_george(golden, name, solver):
    ref, src = _reference_values(name, solver, golden)
    _compare(golden[name + "_lnlike"], ref, _bound(solver, golden[name + "_cond"]), "oracle vs %s (%s) %s" % (solver, src, name))


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("solver", ["BasicSolver", "HODLRSolver"])
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_george(engine, golden, name, solver, exact):
    ref, src = _reference_values(name, solver, golden)
    cs = int(golden[name + "_chunksize"])
    t = golden[name + "_t"]
    engine.unregister_all()
    lc = engine.register_lc(t, golden[name + "_y"], golden[name + "_yerr"], chunks=o.split_chunks(len(t), None if cs < 0 else cs))
    got = engine.lnlike_batch(golden[name + "_theta"], lc_id=lc, exact=exact)
    _compare(got, ref, _bound(solver, golden[name + "_cond"]), "CUDA vs %s (%s) %s exact=%d" % (solver, src, name, exact))


def test_generator_script_is_self_consistent(golden):
    """Without george the script must still import, split chunks like gprot/lc.py:329-342 and cover every golden case."""
    mk = _maker()
    names = sorted({k[: -len("_theta")] for k in golden.files if k.endswith("_theta")})
    assert names == sorted(CASES)
    t = golden["n300c100_t"]
    xs, ys, es = mk.chunk_lists(t, golden["n300c100_y"], golden["n300c100_yerr"], 100)
    ref = o.split_chunks(len(t), 100)
    assert [len(x) for x in xs] == [len(c) for c in ref]
    assert all(np.array_equal(x, t[c]) for x, c in zip(xs, ref))
    assert len(mk.chunk_lists(t, t, t, -1)[0]) == 1
