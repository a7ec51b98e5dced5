#!/usr/bin/env python3
"""Small, fixed target for ncu: a few cfg2 updates (Franka+Ridgeback TrackPoint, K=4096 x T=64)
through the C ABI. Usage: python tools/prof_target.py [cfg2|cfg3|big] [updates]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import engine_lib as el  # noqa: E402
import cases  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if which == "cfg2":
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, 4096, 0.64, precision=abi.FP64, dynamics_mode=abi.DYNAMICS_FUSED)
    e, x0, w = el.Engine(h, abi.default_track_point()), abi.huddled_state(), None
elif which == "cfg3":
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, 16384, 1.28, precision=abi.FP32, dynamics_mode=abi.DYNAMICS_FUSED)
    e, x0, w = el.Engine(h, cases.assisted_params(True, abi.LINKS_BODY_COM)), abi.huddled_state(10.0), cases.constant_wrench(128)
else:
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, 131072, 0.64, precision=abi.FP64, dynamics_mode=abi.DYNAMICS_FUSED)
    e, x0, w = el.Engine(h, abi.default_track_point()), abi.huddled_state(), None
for u in range(n):
    assert e.update(x0, 0.05 * u, w, seed=3) == 0, e.error()
print("ok", which, n, "updates, last device us", e.device_seconds() * 1e6)
e.close()
