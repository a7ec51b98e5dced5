"""Shared case definitions for the controller-level parity tests (oracle vs reference build vs
golden vectors vs CUDA engine)."""
import numpy as np

from assistedmanipulation_b200 import abi


def assisted_params(energy=True, links=abi.LINKS_BODY_COM):
    am = abi.default_assisted_manipulation()
    am.enable_energy_limit = int(energy)
    am.link_position_mode = links
    return am


def constant_wrench(T, force=(10.0, 0.0, 0.0)):
    w = np.zeros((T, 6))
    w[:, :3] = force
    return w


# name -> dict(system, objective, params(), K, horison, keep, threads, x0, updates, cadence, wrench, smoothing)
REF_CASES = {
    "toy_k100_nosmooth": dict(system=abi.SYSTEM_TOY, objective=abi.OBJECTIVE_TOY, params=abi.default_toy_objective, K=100,
                              horison=1.0, keep=0, threads=1, x0=np.zeros(4), updates=6, cadence=0.05, wrench=None, smoothing=None),
    "toy_k253_keep20": dict(system=abi.SYSTEM_TOY, objective=abi.OBJECTIVE_TOY, params=abi.default_toy_objective, K=253,
                            horison=1.0, keep=20, threads=4, x0=np.array([0.1, -0.2, 0.3, 0.0]), updates=6, cadence=0.05,
                            wrench=None, smoothing=(10, 1)),
    "toy_k50_oddcadence": dict(system=abi.SYSTEM_TOY, objective=abi.OBJECTIVE_TOY, params=abi.default_toy_objective, K=50,
                               horison=0.3, keep=20, threads=3, x0=np.zeros(4), updates=8, cadence=0.013, wrench=None,
                               smoothing=(10, 1)),
    "franka_trackpoint_k50": dict(system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_TRACK_POINT,
                                  params=abi.default_track_point, K=50, horison=0.3, keep=20, threads=4,
                                  x0=abi.huddled_state(), updates=5, cadence=0.05, wrench=None, smoothing=(10, 1)),
    "franka_assisted_k60": dict(system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_ASSISTED_MANIPULATION,
                                params=assisted_params, K=60, horison=0.3, keep=20, threads=4, x0=abi.huddled_state(10.0),
                                updates=5, cadence=0.05, wrench=constant_wrench(30), smoothing=(10, 1)),
}

# small cases whose CSV logs, written by the reference's own logger over its own Trajectory, are committed under
# tests/golden/ref_logs/ (tools/gen_log_golden.py)
LOG_CASES = {
    "toy_k6": dict(system=abi.SYSTEM_TOY, objective=abi.OBJECTIVE_TOY, params=abi.default_toy_objective, K=6, horison=0.05, keep=2, threads=1,
                   x0=np.array([0.1, -0.2, 0.3, 0.0]), updates=4, cadence=0.02, wrench=None, smoothing=(2, 1)),
    "franka_trackpoint_k4": dict(system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_TRACK_POINT, params=abi.default_track_point, K=4,
                                 horison=0.03, keep=1, threads=2, x0=abi.huddled_state(), updates=3, cadence=0.01, wrench=None, smoothing=None),
}


def config_for(case, **over):
    kw = dict(keep_best=case["keep"], threads=case["threads"], smoothing=case["smoothing"])
    kw.update(over)
    return abi.make_config(case["system"], case["objective"], case["K"], case["horison"], **kw)


def fp32_costs(ce, co, cost_rtol, max_flips):
    """Single-precision costs against the FP64 oracle on identical inputs. The objectives have steps (1e10 at the inverse
    barriers, cost.hpp:59-61,90-92; 1000 at the hard-coded joint limits, track_point.cpp:48-65): a rollout whose barrier
    argument lands within the fast mode's rounding of a bound at some step takes the step on one side only (a "flip").
    The assisted-manipulation kernels run the whole state path in FP64 (rollout_core.cuh MIXED_SOLVER), so only the
    single-precision kinematics behind the collision-sphere and workspace distances can still flip a rollout: measured
    ~1 in 10^4 rollouts, each worth ~2e-5 of the published control sequence — the caller bounds the count (`max_flips`)
    AND asserts the north-star tolerance on the sequence itself. Every other rollout agrees to `cost_rtol`.
    Returns the number of flipped rollouts."""
    ok = ~np.isnan(co)
    assert np.array_equal(np.isnan(ce), np.isnan(co))
    rel = np.abs(ce[ok] - co[ok]) / np.abs(co[ok])
    flipped = rel > cost_rtol
    assert int(flipped.sum()) <= max_flips, (int(flipped.sum()), max_flips, rel.max())
    assert np.median(rel) <= cost_rtol / 20, np.median(rel)
    return int(flipped.sum())
