"""Pins the oracle's controller restatement (oracle/mppi_oracle.hpp, sg_filter.hpp) against the
REFERENCE ITSELF: src/controller/mppi.cpp, filter.cpp, gaussian.hpp, gram_savitzky_golay.cpp
compiled unmodified (oracle/_ref). Where the built library is absent (GPU box without the
snapshot's _ref) the committed golden vectors generated from it (tools/gen_ref_golden.py) stand in.
Bit-exact: both sides are plain IEEE FP64 compiled with -ffp-contract=off.
"""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle_lib as ol
import ref_lib
from assistedmanipulation_b200 import abi

KEYS = (("optimal", abi.READ_OPTIMAL), ("costs", abi.READ_COSTS), ("weights", abi.READ_WEIGHTS),
        ("gradient", abi.READ_GRADIENT), ("optimal_cost", abi.READ_OPTIMAL_COST))


def run_oracle(oracle, case):
    holder = cases.config_for(case)
    o = ol.Oracle(oracle, holder, case["params"]())
    nu, T, R = holder.cfg.control_dof, o.query(abi.QUERY_STEP_COUNT), o.query(abi.QUERY_ROLLOUT_COUNT)
    sizes = dict(optimal=nu * T, costs=R, weights=R, gradient=nu * T, optimal_cost=1)
    out = {k: [] for k, _ in KEYS}
    out["get"] = []
    for u in range(case["updates"]):
        t = u * case["cadence"]
        assert o.update(case["x0"], t, case["wrench"]) == 0  # GAUSSIAN source: mt19937 like the reference
        for k, what in KEYS:
            out[k].append(o.read(what, sizes[k]))
        out["get"].append(o.get(t + 0.013))
    res = {k: np.stack(v) for k, v in out.items()}
    res["noise_last"] = o.read(abi.READ_NOISE, R * nu * T)
    o.close()
    return res


@pytest.mark.parametrize("name", sorted(cases.REF_CASES))
def test_oracle_matches_reference_golden(oracle, golden, name):
    res = run_oracle(oracle, cases.REF_CASES[name])
    for k, v in res.items():
        g = golden["%s/%s" % (name, k)]
        assert v.shape == g.shape
        assert np.array_equal(v, g), "%s/%s differs from the reference by %g" % (name, k, np.abs(v - g).max())


@pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built here")
@pytest.mark.parametrize("name", ["toy_k253_keep20", "franka_trackpoint_k50"])
def test_golden_is_reproducible_from_reference_build(golden, name):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ol.ROOT, "tools", "gen_ref_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    res = gen.run_reference(ref_lib.load(), cases.REF_CASES[name])
    for k, v in res.items():
        assert np.array_equal(v, golden["%s/%s" % (name, k)])


def test_reference_cannot_exceed_253_rollouts():
    # mppi.hpp:639,642 / mppi.cpp:381: std::uint8_t counters; the driver refuses instead of hanging
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built here")
    ref = ref_lib.load()
    holder = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 254, 1.0)
    p = abi.default_toy_objective()
    assert not ref.ref_create(C.byref(holder.cfg), C.cast(C.byref(p), C.c_void_p), C.sizeof(p))


@pytest.mark.parametrize("key", ["10_0_1_0", "2_0_2_0", "5_5_3_0", "4_0_3_1"])
def test_sg_weights_match_reference(oracle, golden, key):
    m, t, n, s = (int(x) for x in key.split("_"))
    w = np.zeros(2 * m + 1)
    oracle.oracle_sg_weights(m, t, n, s, ol.ptr(w))
    assert np.array_equal(w, golden["sg_weights/" + key])


def test_sg_weights_closed_form(oracle):
    # m=10,n=1 -> 21 x 1/21 ; m=2,n=2 -> [-3,12,17,12,-3]/35
    w = np.zeros(21)
    oracle.oracle_sg_weights(10, 0, 1, 0, ol.ptr(w))
    assert np.allclose(w, 1 / 21, rtol=0, atol=1e-15)
    w = np.zeros(5)
    oracle.oracle_sg_weights(2, 0, 2, 0, ol.ptr(w))
    assert np.allclose(w, np.array([-3, 12, 17, 12, -3]) / 35, rtol=0, atol=1e-15)


def test_sg_window_trace_matches_reference(oracle, golden):
    u, t0s, ref_out = golden["sg_run/u"], golden["sg_run/t0"], golden["sg_run/out"]
    out = np.zeros_like(u)
    u = np.ascontiguousarray(u)
    oracle.oracle_sg_run(u.shape[1], 10, 1, u.shape[0], ol.ptr(np.ascontiguousarray(t0s)), 0.01, ol.ptr(u), ol.ptr(out))
    assert np.array_equal(out, ref_out)


def test_error_conventions(oracle):
    # mppi.cpp:18-69: create returns nullptr + reason
    p = abi.default_toy_objective()

    def fails(**kw):
        h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, kw.pop("K", 8), 1.0, **kw)
        return not oracle.oracle_create(C.byref(h.cfg), C.cast(C.byref(p), C.c_void_p), C.sizeof(p))
    assert fails(K=0)
    assert fails(keep_best=-1)
    assert fails(threads=0)
    assert fails(control_min=np.zeros(3))
    assert fails(covariance=np.eye(3))
    assert not fails()
    # all-NaN rollouts throw (mppi.cpp:368-370)
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 8, 0.1, smoothing=None)
    o = ol.Oracle(oracle, h, p)
    assert o.update(np.array([np.nan, 0, 0, 0]), 0.0, None, np.zeros((10, 10, 2))) == abi.ERR_ALL_NAN
    o.close()
