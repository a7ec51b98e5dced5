"""Worker of test_gpu_variants.test_alternative_builds: runs a few Philox updates of the reach-to-pose controller in
THIS process (the library reads MPPI_B200_BIG_FROM / MPPI_B200_SAMPLE_TILE once, at its first launch) and saves what
the update produced. argv: output.npz precision(0|1) [objective: trackpoint (default) | trackpoint_full | assisted]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import engine_lib as el  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402

import cases  # noqa: E402

out, precision = sys.argv[1], int(sys.argv[2])
which = sys.argv[3] if len(sys.argv) > 3 else "trackpoint"
K, T = 300, 32
objective, params, x0, wrench = abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), abi.huddled_state(), None
if which == "trackpoint_full":
    params.enable_self_collision_avoidance, params.enable_reach_limits, params.link_position_mode = 1, 1, abi.LINKS_BODY_COM
elif which == "assisted":
    objective, params, x0, wrench = abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), abi.huddled_state(10.0), cases.constant_wrench(T)
e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, objective, K, 0.32, keep_best=10, precision=precision, dynamics_mode=abi.DYNAMICS_FUSED), params)
res = {}
for u in range(3):
    assert e.update(x0, 0.05 * u, wrench, seed=11) == 0, e.error()
    res["noise%d" % u] = e.read(abi.READ_NOISE, (K + 2) * T * 12)
    res["costs%d" % u] = e.read(abi.READ_COSTS, K + 2)
    res["U%d" % u] = e.read(abi.READ_OPTIMAL, 12 * T)
    res["weights%d" % u] = e.read(abi.READ_WEIGHTS, K + 2)
    res["gradient%d" % u] = e.read(abi.READ_GRADIENT, 12 * T)
    res["minmax%d" % u] = e.read(abi.READ_MINMAX, 2)
    res["argmin%d" % u] = np.array([e.query(abi.QUERY_ARGMIN)], dtype=np.float64)
e.close()
np.savez(out, **res)
