"""Timing of the assisted-manipulation rollout variants: cfg3 shape (K=16384 x T=128, FP32) and the cfg5 shape
(256 controllers x K=1024 x T=64, FP32, batched), plus FP64 at the cfg3 shape. Prints the device update time."""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import engine_lib as el
import cases
from assistedmanipulation_b200 import abi


def run(label, K, horison, prec, batch=1, n=12):
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, horison, precision=prec, dynamics_mode=abi.DYNAMICS_FUSED, batch=batch)
    e = el.Engine(h, cases.assisted_params(True, abi.LINKS_BODY_COM))
    T = e.query(abi.QUERY_STEP_COUNT)
    x0 = np.tile(abi.huddled_state(10.0), (batch, 1))
    w = np.tile(cases.constant_wrench(T), (batch, 1, 1))
    ds = []
    for u in range(n):
        assert e.update(x0, 0.05 * u, w, seed=5) == 0, e.error()
        ds.append(e.device_seconds())
    print(label, 'K', K, 'T', T, 'batch', batch, 'device us %.1f' % (np.median(ds[3:]) * 1e6), flush=True)
    e.close()


run('cfg3 f32', 16384, 1.28, abi.FP32)
run('cfg3 f64', 16384, 1.28, abi.FP64)
run('cfg5 f32', 1024, 0.64, abi.FP32, batch=256)
run('cfg2-shape AM f64', 4096, 0.64, abi.FP64)
