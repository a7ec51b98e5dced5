"""The C++ facade (assistedmanipulation_b200/cpp: mppi::Trajectory / Dynamics / Cost with the
reference's API, src/controller/mppi.hpp:267-474) driven like the reference's Actor, compared with the
oracle on identical injected noise."""
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi
from test_abi_cpu import build_facade_demo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["toy", "track", "assisted", "assisted_locf"])
def test_facade_matches_oracle(oracle, tmp_path, which):
    exe = build_facade_demo()
    K, horison, updates = 126, 0.3, 4
    T = 30
    if which == "toy":
        system, objective, params, nu, x0, wrench, sigma = abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective(), 2, np.zeros(4), None, np.ones(2)
    else:
        system, nu, x0, sigma = abi.SYSTEM_FRANKA_RIDGEBACK, 12, abi.huddled_state(10.0), np.sqrt(abi.FRANKA_COVARIANCE_DIAG)
        if which == "track":
            objective, params, wrench = abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), None
        else:
            objective, params, wrench = abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), cases.constant_wrench(T)
    R = K + 2
    noise = np.random.default_rng(1).standard_normal((updates, R, T, nu)) * sigma
    noise.tofile(str(tmp_path / "noise.bin"))
    r = subprocess.run([exe, which, str(K), str(horison), str(updates), str(tmp_path / "noise.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "foreign_cost_refused 1" in r.stdout and "no device implementation" in r.stderr
    rec = np.fromfile(str(tmp_path / "out.bin")).reshape(updates, nu * T + nu + 1 + R)
    holder = abi.make_config(system, objective, K, horison, keep_best=20, threads=4, control_default=np.zeros(nu) if which != "toy" else None)
    o = ol.Oracle(oracle, holder, params)
    for u in range(updates):
        assert o.update(x0, 0.05 * u, wrench, noise[u]) == 0
        Uo = o.read(abi.READ_OPTIMAL, nu * T)
        assert np.abs(rec[u, :nu * T] - Uo).max() <= 1e-9 * np.abs(Uo).max()
        assert np.allclose(rec[u, nu * T:nu * T + nu], o.get(0.05 * u + 0.013), rtol=1e-9, atol=1e-9 * np.abs(Uo).max())
        oc = o.read(abi.READ_OPTIMAL_COST, 1)[0]
        assert abs(rec[u, nu * T + nu] - oc) <= 1e-8 * abs(oc)
        assert np.allclose(rec[u, nu * T + nu + 1:], o.read(abi.READ_WEIGHTS, R), rtol=1e-8, atol=1e-14)
    if which.startswith("assisted"):
        got = np.array([float(x) for x in r.stdout.split("breakdown")[1].split("\n")[0].split()])
        assert np.allclose(got, o.read(abi.READ_BREAKDOWN, 8)[:7], rtol=1e-8, atol=1e-8)
    o.close()


GOLDEN_ARGS = {
    # case -> (which, K, horison, updates, keep, smoothing, cadence, extra x0 args)
    "toy_k100_nosmooth": ("toy", 100, 1.0, 6, 0, 0, 0.05, ["0", "0", "0", "0"]),
    "toy_k253_keep20": ("toy", 253, 1.0, 6, 20, 1, 0.05, ["0.1", "-0.2", "0.3", "0.0"]),
    "toy_k50_oddcadence": ("toy", 50, 0.3, 8, 20, 1, 0.013, ["0", "0", "0", "0"]),
    "franka_trackpoint_k50": ("track", 50, 0.3, 5, 20, 1, 0.05, ["100.0"]),
    "franka_assisted_k60": ("assisted", 60, 0.3, 5, 20, 1, 0.05, ["10.0"]),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_ARGS))
def test_reference_rng_mode_reproduces_the_reference_build(golden, tmp_path, name):
    """End to end against the REFERENCE ITSELF: tests/golden/ref_mppi.npz holds what the reference's own
    mppi.cpp / filter.cpp / gaussian.hpp (compiled unmodified, oracle/_ref) published over several closed-loop
    updates with its mt19937 sampling. The facade in reference-RNG mode draws the same stream on the host,
    the device does everything else; the control sequences must agree to 1e-9."""
    which, K, horison, updates, keep, smoothing, cadence, extra = GOLDEN_ARGS[name]
    exe = build_facade_demo()
    r = subprocess.run([exe, which, str(K), str(horison), str(updates), "-", str(tmp_path / "out.bin"), str(keep), str(smoothing), str(cadence), "refrng"] + extra,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    nu = 2 if which == "toy" else 12
    T, R = int(np.ceil(horison / 0.01)), K + 2
    rec = np.fromfile(str(tmp_path / "out.bin")).reshape(updates, nu * T + nu + 1 + R)
    U_ref, w_ref, oc_ref, get_ref = golden[name + "/optimal"], golden[name + "/weights"], golden[name + "/optimal_cost"], golden[name + "/get"]
    for u in range(updates):
        scale = np.abs(U_ref[u]).max()
        assert np.abs(rec[u, :nu * T] - U_ref[u]).max() <= 1e-9 * scale, (u, np.abs(rec[u, :nu * T] - U_ref[u]).max() / scale)
        assert np.allclose(rec[u, nu * T:nu * T + nu], get_ref[u], rtol=1e-9, atol=1e-9 * scale)
        assert abs(rec[u, nu * T + nu] - oc_ref[u, 0]) <= 1e-8 * abs(oc_ref[u, 0])
        assert np.allclose(rec[u, nu * T + nu + 1:], w_ref[u], rtol=1e-8, atol=1e-14)


LOG_ARGS = {
    # case -> (which, K, horison, updates, keep, smoothing, cadence, x0 args, SG window)
    "toy_k6": ("toy", 6, 0.05, 4, 2, 1, 0.02, ["0.1", "-0.2", "0.3", "0.0"], 2),
    "franka_trackpoint_k4": ("track", 4, 0.03, 3, 1, 0, 0.01, ["100.0"], 10),
}


@pytest.mark.parametrize("name", sorted(LOG_ARGS))
def test_cpp_logger_writes_the_reference_files(tmp_path, name):
    """logger::MPPI of the facade (cpp/mppi_b200/logging.hpp) over the device trajectory in
    reference-RNG mode against the files the reference's own logger wrote for its own Trajectory
    (tests/golden/ref_logs): same files, same headers, same row keys; values to the 6 printed digits."""
    from assistedmanipulation_b200 import formats
    exe = build_facade_demo()
    which, K, horison, updates, keep, smoothing, cadence, extra, window = LOG_ARGS[name]
    folder = tmp_path / "logs"
    env = dict(os.environ, MPPI_B200_DEMO_LOG=str(folder), MPPI_B200_DEMO_SG_WINDOW=str(window))
    r = subprocess.run([exe, which, str(K), str(horison), str(updates), "-", str(tmp_path / "out.bin"), str(keep), str(smoothing), str(cadence), "refrng"] + extra,
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    golden = os.path.join(ol.ROOT, "tests", "golden", "ref_logs", name)
    assert sorted(os.listdir(folder)) == sorted(os.listdir(golden))
    identical = 0
    for f in sorted(os.listdir(golden)):
        got_text, want_text = open(folder / f).read(), open(os.path.join(golden, f)).read()
        assert got_text.splitlines()[0] == want_text.splitlines()[0]          # header, byte for byte
        (_, got), (_, want) = formats.read_csv(str(folder / f)), formats.read_csv(os.path.join(golden, f))
        assert got.shape == want.shape, f
        cols = slice(0, 2) if f == "update.csv" else slice(None)                # the third column of update.csv is a duration
        np.testing.assert_allclose(got[:, cols], want[:, cols], rtol=2e-5, atol=1e-9, err_msg=f)
        identical += got_text == want_text
    assert identical >= 3, identical   # most files agree to the last printed digit
