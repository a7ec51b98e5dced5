#!/usr/bin/env python3
"""Small, fixed target for ncu: a few updates of one BASELINE.json workload through the C ABI.
Usage: python tools/prof_target.py [cfg2|cfg2_f32|cfg3|cfg4_f64|cfg4_f32] [updates]   (cfg4_*: one GPU's shard of 8, K = 131072)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import cases  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402
from assistedmanipulation_b200 import engine as el  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
TP = (abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), abi.huddled_state(), None)
AM = (abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), abi.huddled_state(10.0), cases.constant_wrench(128))
objective, params, x0, w, K, hor, prec = {
    "cfg2": TP + (4096, 0.64, abi.FP64), "cfg2_f32": TP + (4096, 0.64, abi.FP32), "cfg3": AM + (16384, 1.28, abi.FP32),
    "cfg4_f64": TP + (131072, 0.64, abi.FP64), "cfg4_f32": TP + (131072, 0.64, abi.FP32), "big": TP + (131072, 0.64, abi.FP64)}[which]
e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, objective, K, hor, precision=prec, dynamics_mode=abi.DYNAMICS_FUSED), params)
for u in range(n):
    assert e.update(x0, 0.05 * u, w, seed=3) == 0, e.error()
print("ok", which, n, "updates, last device us", e.device_seconds() * 1e6, "optimal cost", e.read(abi.READ_OPTIMAL_COST, 1)[0])
e.close()
