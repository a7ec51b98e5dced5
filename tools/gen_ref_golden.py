#!/usr/bin/env python3
"""Generate tests/golden/ref_mppi.npz from oracle/_ref/libmppi_ref.so — the REFERENCE's own
mppi.cpp / filter.cpp / gaussian.hpp / gram_savitzky_golay.cpp compiled unmodified from
/root/reference (oracle/Makefile `ref`). Run in the build container (needs /root/reference):

    make -C oracle ref && python tools/gen_ref_golden.py

For every case of tests/cases.py::REF_CASES the reference runs `updates` consecutive
Trajectory::update calls with its own mt19937 Gaussian sampling; after each update we store the
published trajectory, per-rollout costs, weights, gradient, the optimal re-rollout cost and one
Trajectory::get readout. The noise of the LAST update is stored as well.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import ref_lib  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402

_dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def run_reference(ref, case):
    holder = cases.config_for(case)
    params = case["params"]()
    h = ref.ref_create(C.byref(holder.cfg), C.cast(C.byref(params), C.c_void_p), C.sizeof(params))
    assert h
    nu = holder.cfg.control_dof
    T = int(np.ceil(case["horison"] / 0.01))
    R = case["K"] + 2
    out = {k: [] for k in ("optimal", "costs", "weights", "gradient", "optimal_cost", "get")}
    x0 = np.ascontiguousarray(case["x0"], dtype=np.float64)
    w = None if case["wrench"] is None else np.ascontiguousarray(case["wrench"])
    for u in range(case["updates"]):
        t = u * case["cadence"]
        assert ref.ref_update(h, ptr(x0), t, ptr(w)) == 0
        for key, what, n in (("optimal", abi.READ_OPTIMAL, nu * T), ("costs", abi.READ_COSTS, R), ("weights", abi.READ_WEIGHTS, R),
                             ("gradient", abi.READ_GRADIENT, nu * T), ("optimal_cost", abi.READ_OPTIMAL_COST, 1)):
            b = np.zeros(n)
            assert ref.ref_read(h, what, ptr(b), b.nbytes) == 0
            out[key].append(b)
        g = np.zeros(nu)
        ref.ref_get(h, ptr(g), t + 0.013)
        out["get"].append(g)
    noise = np.zeros(R * nu * T)
    assert ref.ref_read(h, abi.READ_NOISE, ptr(noise), noise.nbytes) == 0
    ref.ref_destroy(h)
    res = {k: np.stack(v) for k, v in out.items()}
    res["noise_last"] = noise
    return res


def main():
    assert ref_lib.available(), "build oracle/_ref first: make -C oracle ref"
    ref = ref_lib.load()
    blob = {}
    for name, case in cases.REF_CASES.items():
        for k, v in run_reference(ref, case).items():
            blob["%s/%s" % (name, k)] = v
    # Savitzky-Golay weights and a stateful window trace straight from the reference sources
    for (m, t, n, s) in ((10, 0, 1, 0), (2, 0, 2, 0), (5, 5, 3, 0), (4, 0, 3, 1)):
        w = np.zeros(2 * m + 1)
        ref.ref_sg_weights(m, t, n, s, ptr(w))
        blob["sg_weights/%d_%d_%d_%d" % (m, t, n, s)] = w
    rng = np.random.default_rng(7)
    steps, window, updates = 30, 10, 6
    u = rng.standard_normal((updates, steps))
    t0s = np.arange(updates) * 0.05
    out = np.zeros_like(u)
    ref.ref_sg_run(steps, window, 1, updates, ptr(t0s), 0.01, ptr(u), ptr(out))
    blob["sg_run/u"], blob["sg_run/t0"], blob["sg_run/out"] = u, t0s, out
    path = os.path.join(ROOT, "tests", "golden", "ref_mppi.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
