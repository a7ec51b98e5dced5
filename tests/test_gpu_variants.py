"""Every kernel variant (systems x objectives x FP64/FP32 x FAITHFUL/FUSED x single/batched) runs a few
updates with keep-best, smoothing, Philox and injected noise without error and publishes finite controls
(compute-sanitizer is closed on this pool; this is the broad exercise, parity is in test_gpu_parity.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu

RUNS = [
    ("toy", abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective, 70, 0.24, lambda: np.zeros(4), None, 2, 1),
    ("track", abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point, 70, 0.2, abi.huddled_state, None, 12, 1),
    ("assisted", abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, lambda: cases.assisted_params(True, abi.LINKS_BODY_COM), 70, 0.2,
     lambda: abi.huddled_state(10.0), cases.constant_wrench(20), 12, 1),
    ("assisted_batch3", abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, lambda: cases.assisted_params(False, abi.LINKS_ZERO), 40, 0.2,
     lambda: abi.huddled_state(10.0), cases.constant_wrench(20), 12, 3),
]


@pytest.mark.parametrize("run", RUNS, ids=[r[0] for r in RUNS])
@pytest.mark.parametrize("precision", [abi.FP64, abi.FP32])
@pytest.mark.parametrize("mode", [abi.DYNAMICS_FUSED, abi.DYNAMICS_FAITHFUL])
def test_variant_runs(run, precision, mode):
    import engine_lib as el
    _, system, objective, params, K, horison, x0, w, nu, batch = run
    e = el.Engine(abi.make_config(system, objective, K, horison, keep_best=8, precision=precision, dynamics_mode=mode, batch=batch), params())
    T = e.query(abi.QUERY_STEP_COUNT)
    xs = np.stack([x0()] * batch)
    ws = None if w is None else np.stack([w] * batch)
    rng = np.random.default_rng(0)
    for u in range(3):
        noise = rng.standard_normal((batch, K + 2, T, nu)) if u == 1 else None
        assert e.update(xs, 0.05 * u, ws, noise, seed=1) == 0, e.error()
        U = e.read(abi.READ_OPTIMAL, batch * nu * T)
        assert np.all(np.isfinite(U))
    assert np.all(np.isfinite(e.read(abi.READ_OPTIMAL_COST, batch)))
    e.close()


def _alt_run(tmp_path, name, precision, env, objective="trackpoint"):
    out = str(tmp_path / (name + ".npz"))
    full = dict(os.environ)
    full.update(env)
    subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "alt_worker.py"), out, str(precision), objective], check=True, env=full, timeout=300)
    return np.load(out)


@pytest.mark.parametrize("objective", ["assisted", "trackpoint_full"])
def test_two_warp_rollout_kernel_is_bit_identical_with_the_one_warp_path(tmp_path, objective):
    """FP32 fast mode of the objectives with kinematics: the production kernel splits a rollout over a state warp (FP64 state
    path) and a cost warp (FP32 kinematics / RNEA / objective) that meet in a shared-memory ring (k_rollout.cuh
    k_rollout_split); MPPI_B200_SPLIT=0 runs the same arithmetic in one warp (rollout_franka, the code the CPU tests
    compile). Same operations in the same order: every cost, and therefore everything downstream, is bit-identical —
    over updates with a kept set and a time shift."""
    split = _alt_run(tmp_path, "split", abi.FP32, {}, objective)
    one = _alt_run(tmp_path, "one_warp", abi.FP32, {"MPPI_B200_SPLIT": "0"}, objective)
    for k in split.files:
        assert np.array_equal(split[k], one[k], equal_nan=True), k


@pytest.mark.parametrize("precision", [abi.FP64, abi.FP32])
@pytest.mark.parametrize("objective", ["trackpoint", "assisted"])
def test_fused_tail_of_the_update_matches_the_separate_kernels(tmp_path, precision, objective):
    """One rank with a small rollout set: the weighted-sum kernel computes the weights of its own rollouts and k_finish
    combines the partial sums (k_misc.cu, FUSED). MPPI_B200_FUSED_TAIL=0 runs k_weights / k_gradient / k_gradient_reduce /
    k_finish as separate kernels — what a sharded or a large rollout set always runs. Same costs, minimum and best rollout
    bit for bit; the sums are taken in another (fixed) order, so the normalised weights, the gradient and the controls agree
    to rounding — over updates with a kept set and a shift."""
    fused = _alt_run(tmp_path, "fused", precision, {}, objective)
    apart = _alt_run(tmp_path, "apart", precision, {"MPPI_B200_FUSED_TAIL": "0"}, objective)
    for k in ("noise", "costs", "minmax", "argmin"):
        assert np.array_equal(fused[k + "0"], apart[k + "0"], equal_nan=True), k
    for u in range(3):   # (the rollouts of update u > 0 start from controls that differ in the last bits)
        for k in ("weights", "gradient", "U"):
            a, b = fused[k + str(u)], apart[k + str(u)]
            assert np.abs(a - b).max() <= (1e-12 if u == 0 else 1e-8) * np.abs(b).max(), (k, u)


@pytest.mark.parametrize("precision", [abi.FP64, abi.FP32])
def test_rollout_blocks_that_draw_their_own_noise_are_bit_identical_with_the_sampling_kernel(tmp_path, precision):
    """Small rollout sets of the reach-to-pose controller (a grid of one-warp blocks no larger than the machine): seven more
    warps per rollout block draw the block's noise chunk by chunk while warp 0 integrates (rollout_core.cuh "noise chase");
    MPPI_B200_CHASE=0 runs k_sample_columns ahead of the rollout kernel as for every other configuration. Same counters,
    same arithmetic, same order of every sum: all buffers are bit-identical over updates with a kept set and a time shift."""
    chase = _alt_run(tmp_path, "chase", precision, {})
    apart = _alt_run(tmp_path, "sampling_kernel", precision, {"MPPI_B200_CHASE": "0"})
    for k in chase.files:
        assert np.array_equal(chase[k], apart[k], equal_nan=True), k


@pytest.mark.parametrize("precision,c_tol,u_tol", [(abi.FP64, 1e-9, 1e-9), (abi.FP32, 1e-3, 2e-4)])
def test_alternative_builds(tmp_path, precision, c_tol, u_tol):
    """The builds kept for A/B runs against the defaults (unrolled rollout kernel, k_sample_quads), each in its own
    process because the library reads the switches once: the tile sampling kernel (MPPI_B200_SAMPLE_TILE=1) must
    reproduce every buffer bit for bit over updates with a kept set and a shift (same counters, same arithmetic, same
    rollout kernel); the loop-body rollout kernel (MPPI_B200_BIG_FROM above the rollout count) sees identical noise
    and must agree on costs and controls within the arithmetic's tolerance."""
    base = _alt_run(tmp_path, "default", precision, {})
    tile = _alt_run(tmp_path, "tile", precision, {"MPPI_B200_SAMPLE_TILE": "1"})
    loop = _alt_run(tmp_path, "loop_body", precision, {"MPPI_B200_BIG_FROM": "1000000000"})
    for k in base.files:
        assert np.array_equal(base[k], tile[k]), k
    assert np.array_equal(base["noise0"], loop["noise0"])
    rel = np.abs(loop["costs0"] - base["costs0"]) / np.abs(base["costs0"])
    if precision == abi.FP64:
        assert rel.max() <= c_tol
    else:
        # single precision: a joint that lands within rounding of one of the hard-coded limits at some step takes the
        # 1000-unit penalty in one build and not in the other (track_point.cpp:48-65) — allowed for a handful of rollouts
        assert (rel <= c_tol).mean() >= 0.98
    if rel.max() <= c_tol:
        assert np.abs(loop["U0"] - base["U0"]).max() <= u_tol * np.abs(base["U0"]).max()
