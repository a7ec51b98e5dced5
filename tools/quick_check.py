"""One short GPU call: smoke() (parity against the oracle) and device timings of the three rollout kernels (torch-free)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
t0 = time.time()
ONLY_CFG2 = '--cfg2-only' in sys.argv
if not ONLY_CFG2:
    import __graft_entry__ as g
    g.smoke()
    print('smoke %.1f s' % (time.time() - t0), flush=True)
import engine_lib as el
import cases
from assistedmanipulation_b200 import abi
import ctypes as C
def run(label, obj, params, K, hor, prec, x0, wrench=None, n=40):
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, obj, K, hor, precision=prec, dynamics_mode=abi.DYNAMICS_FUSED)
    e = el.Engine(h, params)
    e.lib.mppi_b200_set_profiling(e.h, 1)
    st = np.zeros((n, 8)); ds = []
    for u in range(n):
        assert e.update(x0, 0.05 * u, wrench, seed=5) == 0, e.error()
        e.lib.mppi_b200_stage_seconds(e.h, st[u].ctypes.data_as(C.POINTER(C.c_double)), 8); ds.append(e.device_seconds())
    m = np.median(st[5:], axis=0) * 1e6
    print(label, 'K', K, 'device us %.1f' % (np.median(ds[5:]) * 1e6), ' '.join('%s=%.1f' % (a, b) for a, b in zip(abi.STAGES, m)), flush=True)
    e.lib.mppi_b200_set_profiling(e.h, 0)
    ds = []
    for u in range(n, 2 * n):
        assert e.update(x0, 0.05 * u, wrench, seed=5) == 0, e.error()
        ds.append(e.device_seconds())
    print(label, 'graph replay device us p50 %.1f' % (np.median(ds[3:]) * 1e6), flush=True)
    e.close()
x = abi.huddled_state()
run('cfg2 f64', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 4096, 0.64, abi.FP64, x)
if ONLY_CFG2:
    print('total %.1f s' % (time.time() - t0), flush=True)
    sys.exit(0)
run('big f64', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 131072, 0.64, abi.FP64, x, n=10)
run('cfg3 f32', abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, 1), 16384, 1.28, abi.FP32, abi.huddled_state(10.0), cases.constant_wrench(128), n=12)
run('cfg2 f32', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 4096, 0.64, abi.FP32, x)
print('total %.1f s' % (time.time() - t0), flush=True)
