"""Pins the oracle's objectives, cost functors, energy tank and state presets (oracle/systems.hpp) against the
REFERENCE ITSELF: src/frankaridgeback/objective/track_point.cpp, objective/assisted_manipulation.cpp,
controller/cost.hpp, controller/energy.hpp and frankaridgeback/state.cpp compiled unmodified into
oracle/_ref/libmppi_ref.so (oracle/Makefile `ref`) and driven through a FrankaRidgeback::Dynamics that serves the
probe's kinematics. Where the built library is absent (GPU box) the committed golden vectors generated from it
(tools/gen_objective_golden.py -> tests/golden/ref_objective.npz) stand in.
Bit-exact: both sides are plain IEEE FP64 compiled with -ffp-contract=off, same order of operations."""
import ctypes as C
import os

import numpy as np
import pytest

import objective_probe as op
import oracle_lib as ol
import ref_lib
from assistedmanipulation_b200 import abi

_dp = C.POINTER(C.c_double)
GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_objective.npz"))
TERMS = ("total", "joint", "self_collision", "workspace", "energy", "velocity", "trajectory", "manipulability")


def _records():
    return op.records(int(GOLDEN["records_seed"]), int(GOLDEN["records_count"]))


@pytest.mark.parametrize("name", sorted(op.variants()))
def test_oracle_objective_matches_the_reference_term_by_term(oracle, name):
    objective, params = op.variants()[name]
    got = op.evaluate(oracle.oracle_objective_probe, objective, params, _records())
    ref = GOLDEN["probe/" + name]
    assert not np.isnan(ref).any()
    for i, term in enumerate(TERMS):
        assert np.array_equal(got[:, i], ref[:, i]), "%s/%s differs from the reference by %g" % (name, term, np.abs(got[:, i] - ref[:, i]).max())
    # the probe set exercises the jumps and the plain branches of every term that is on
    if objective == abi.OBJECTIVE_ASSISTED_MANIPULATION:
        for i in (1, 2, 3, 7):
            assert (ref[:, i] >= 1e10).any() and (ref[:, i] < 1e10).any(), term
        assert (ref[:, 6] == 0).any() and (ref[:, 6] > 0).any()


@pytest.mark.parametrize("name", sorted(op.FUNCTORS))
def test_cost_functors_match_the_reference(oracle, name):
    kind, a, b, c, d = op.FUNCTORS[name]
    v = op.functor_values(kind, a)
    fn = {0: lambda x: oracle.oracle_quadratic(a, b, c, x), 1: lambda x: oracle.oracle_left_barrier(a, b, c, x),
          2: lambda x: oracle.oracle_right_barrier(a, b, c, x), 3: lambda x: oracle.oracle_upper_log_barrier(a, b, c, d, x),
          4: lambda x: oracle.oracle_lower_log_barrier(a, b, c, d, x)}[kind]
    got = np.array([fn(float(x)) for x in v])
    assert np.array_equal(got, GOLDEN["functor/" + name])


def test_energy_tank_matches_the_reference(oracle):
    power, ref = GOLDEN["tank/power"], GOLDEN["tank/energy"]
    e, out = 10.0, np.zeros(2)
    for p, r in zip(power, ref):
        oracle.oracle_tank(e, float(p), 0.01, out.ctypes.data_as(_dp))
        e = out[0]
        assert e == r and out[1] == np.sqrt(2.0 * r)
    assert (ref == 0.0).any() and (ref > 0.0).any()   # the max(0, .) floor is visited


def test_state_presets_match_the_reference():
    presets = GOLDEN["presets"]
    assert np.array_equal(presets[1], abi.huddled_state())         # Preset::HUDDLED, state.cpp:15-19
    assert np.array_equal(presets[0], np.zeros(31))                # Preset::ZERO returns before the energy is set
    assert (presets[1:, 30] == 100.0).all()


@pytest.mark.parametrize("case", ["kalman1", "locf", "average"])
def test_dynamics_forecast_oracle_matches_the_reference(oracle, case):
    """the reference's own DynamicsForecast (frankaridgeback/dynamics.cpp) and wrench forecasters over the
    oracle's rigid-body core against the oracle's restatement of the loop: recording order, wrench lookups, indexing."""
    import dynamics_forecast_script as dfs
    name, typ, hw, fdt, order = [c for c in dfs.CASES if c[0] == case][0]
    f = dfs.Oracle(oracle, typ, hw, fdt, order)
    rec, idx = dfs.run_session(f, 31)
    f.close()
    want = GOLDEN["dynamics_forecast/%s/records" % case]
    assert rec.shape == want.shape == (3, 20, abi.DYNAMICS_FORECAST_RECORD)
    assert np.array_equal(rec, want), np.abs(rec - want).max()
    assert np.array_equal(idx, GOLDEN["dynamics_forecast/%s/parameterise" % case])
    assert np.abs(want[..., abi.DF_WRENCH]).max() > 1.0


@pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built here")
def test_objective_golden_is_reproducible_from_the_reference_build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_objective", os.path.join(ol.ROOT, "tools", "gen_objective_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = gen.reference_vectors(ref_lib.load())
    assert sorted(out) == sorted(GOLDEN.files)
    for k, v in out.items():
        assert np.array_equal(v, GOLDEN[k]), k
