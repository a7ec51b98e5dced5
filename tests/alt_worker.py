"""Worker of test_gpu_variants.test_alternative_builds: runs a few Philox updates of the reach-to-pose controller in
THIS process (the library reads MPPI_B200_BIG_FROM / MPPI_B200_SAMPLE_TILE once, at its first launch) and saves what
the update produced. argv: output.npz precision(0|1)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import engine_lib as el  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402

out, precision = sys.argv[1], int(sys.argv[2])
K, T = 300, 32
e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, keep_best=10, precision=precision, dynamics_mode=abi.DYNAMICS_FUSED),
              abi.default_track_point())
x0 = abi.huddled_state()
res = {}
for u in range(3):
    assert e.update(x0, 0.05 * u, seed=11) == 0, e.error()
    res["noise%d" % u] = e.read(abi.READ_NOISE, (K + 2) * T * 12)
    res["costs%d" % u] = e.read(abi.READ_COSTS, K + 2)
    res["U%d" % u] = e.read(abi.READ_OPTIMAL, 12 * T)
e.close()
np.savez(out, **res)
