"""One short GPU call: smoke() (parity against the oracle) and device timings of the rollout kernels (torch-free).
  python tools/quick_check.py [--no-smoke] [--only cfg2,big,cfg3,cfg2f32,cfg5] [--flips]
Kernel switches are environment variables read by the library (MPPI_B200_LIB, MPPI_B200_AM_BLOCK, MPPI_B200_SPLIT, ...)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
t0 = time.time()
args = sys.argv[1:]
only = None
for i, a in enumerate(args):
    if a == '--only':
        only = args[i + 1].split(',')
if '--no-smoke' not in args and '--cfg2-only' not in args:
    import __graft_entry__ as g
    g.smoke()
    print('smoke %.1f s' % (time.time() - t0), flush=True)
if '--cfg2-only' in args:
    only = ['cfg2']
from assistedmanipulation_b200 import engine as el
import cases
from assistedmanipulation_b200 import abi
import ctypes as C
def run(label, obj, params, K, hor, prec, x0, wrench=None, n=40, batch=1):
    if only is not None and label.split()[0] not in only:
        return
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, obj, K, hor, precision=prec, dynamics_mode=abi.DYNAMICS_FUSED, batch=batch)
    e = el.Engine(h, params)
    if batch > 1:
        x0 = np.ascontiguousarray(np.tile(x0, (batch, 1)))
        wrench = None if wrench is None else np.ascontiguousarray(np.tile(wrench, (batch, 1, 1)))
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
def flips():
    """FP32 fast mode at full config-3 size against the FP64 oracle on the same noise: flipped rollouts and the error of U"""
    import oracle_lib as ol
    K, T, nu = 16384, 128, 12
    params, W, x0 = cases.assisted_params(True, abi.LINKS_BODY_COM), cases.constant_wrench(T), abi.huddled_state(10.0)
    o = ol.Oracle(ol.load(), abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 1.28, threads=16, smoothing=None, control_bound=False), params)
    for seed in (3, 4, 5):
        e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 1.28, precision=abi.FP32, dynamics_mode=abi.DYNAMICS_FUSED,
                                      smoothing=None, control_bound=False), params)
        assert e.update(x0, 0.0, W, seed=seed) == 0, e.error()
        noise = e.read(abi.READ_NOISE, (K + 2) * T * nu)
        o2 = ol.Oracle(ol.load(), abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 1.28, threads=16, smoothing=None, control_bound=False), params)
        assert o2.update(x0, 0.0, W, noise) == 0
        Uo, Ue = o2.read(abi.READ_OPTIMAL, nu * T), e.read(abi.READ_OPTIMAL, nu * T)
        co, ce = o2.read(abi.READ_COSTS, K + 2), e.read(abi.READ_COSTS, K + 2)
        fl = np.abs(ce - co) > 5e9
        print('flips seed', seed, int(fl.sum()), 'U err %.2e' % (np.abs(Ue - Uo).max() / np.abs(Uo).max()), 'median cost rel %.2e' % np.median(np.abs(ce - co) / np.abs(co)), flush=True)
        e.close(); o2.close()
    o.close()
x = abi.huddled_state()
run('cfg2 f64', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 4096, 0.64, abi.FP64, x)
run('big f64', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 131072, 0.64, abi.FP64, x, n=10)
run('bigf32 f32', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 131072, 0.64, abi.FP32, x, n=10)
run('cfg3 f32', abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, 1), 16384, 1.28, abi.FP32, abi.huddled_state(10.0), cases.constant_wrench(128), n=12)
run('cfg2f32 f32', abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 4096, 0.64, abi.FP32, x)
run('cfg5 f32 batch32', abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, 1), 1024, 0.64, abi.FP32, abi.huddled_state(10.0), cases.constant_wrench(64), n=12, batch=32)
if '--flips' in args:
    flips()
print('total %.1f s' % (time.time() - t0), flush=True)
