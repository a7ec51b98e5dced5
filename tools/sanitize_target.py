#!/usr/bin/env python3
"""Small multi-feature run for compute-sanitizer: toy + Franka (both objectives, both precisions, both
dynamics modes), keep-best, smoothing, injected and Philox noise, a batched engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from assistedmanipulation_b200 import engine as el  # noqa: E402
import cases  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402

rng = np.random.default_rng(0)
runs = [
    (abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective(), 70, 0.24, np.zeros(4), None, 2, 1),
    (abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 70, 0.2, abi.huddled_state(), None, 12, 1),
    (abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), 70, 0.2, abi.huddled_state(10.0), cases.constant_wrench(20), 12, 1),
    (abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), 40, 0.2, abi.huddled_state(10.0), cases.constant_wrench(20), 12, 3),
]
for system, objective, params, K, horison, x0, w, nu, batch in runs:
    for precision in (abi.FP64, abi.FP32):
        for mode in (abi.DYNAMICS_FUSED, abi.DYNAMICS_FAITHFUL):
            h = abi.make_config(system, objective, K, horison, keep_best=8, precision=precision, dynamics_mode=mode, batch=batch)
            e = el.Engine(h, params)
            T = e.query(abi.QUERY_STEP_COUNT)
            xs = np.stack([x0] * batch)
            ws = None if w is None else np.stack([w] * batch)
            for u in range(3):
                noise = rng.standard_normal((batch, K + 2, T, nu)) if u == 1 else None
                assert e.update(xs, 0.05 * u, ws, noise, seed=1) == 0, e.error()
            e.read(abi.READ_OPTIMAL, batch * nu * T)
            e.close()
print("sanitize target ok")
