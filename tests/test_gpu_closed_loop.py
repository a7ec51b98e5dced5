"""Closed loop on the toy system of BASELINE.json config 1: the CUDA controller (C ABI, in-kernel Philox noise,
CUDA-graph replay, warm start, smoothing) drives a double integrator to the target the way Actor::act drives the
robot (controller update every 0.05 s, Trajectory::get every simulation step, actor.cpp:168-201). A functional
check that the update is a controller and not only a parity artefact. (The Franka+Ridgeback objectives are tuned
for the reference's RaiSim plant, whose arm takes velocity commands; closing that loop over the torque-driven
rollout model says nothing about this engine, so it is not asserted here.)"""
import numpy as np
import pytest

import engine_lib as el
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", [abi.FP64, abi.FP32])
def test_double_integrator_reaches_the_target(precision):
    params = abi.default_toy_objective()
    target = np.array(list(params.target))
    holder = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 1024, 1.0, precision=precision, keep_best=32)
    e = el.Engine(holder, params)
    x = np.array([-0.5, 0.25, 0.0, 0.0])           # px, py, vx, vy
    d0 = np.linalg.norm(x[:2] - target)
    dt, t = 0.01, 0.0
    distances = []
    for update in range(80):
        assert e.update(x, t, None, seed=3) == 0, e.error()
        for _ in range(5):
            u = e.get(t)
            assert np.isfinite(u).all() and np.all(np.abs(u) <= 5.0 + 1e-12)   # control_min / control_max
            x[2:] += u * dt                        # the toy dynamics of the rollout kernel (semi-implicit Euler)
            x[:2] += x[2:] * dt
            t += dt
        distances.append(np.linalg.norm(x[:2] - target))
    e.close()
    assert distances[-1] < 0.1 * d0, (d0, distances[::10])
    assert np.linalg.norm(x[2:]) < 0.5                                           # and came to rest there
    assert max(distances[40:]) < 0.25 * d0                                       # without leaving again
