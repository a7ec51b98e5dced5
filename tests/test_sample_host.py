"""The sampling kernel's per-thread core (csrc/sample_core.cuh) compiled for the CPU: Philox4x32-10 against the
Random123 known-answer vectors, and the quad enumeration of k_sample_quads — which rollout, step, counter and element
every thread of the grid produces — against an independent numpy construction of the noise buffer (static rollouts
mppi.cpp:222 / :269, kept rollouts untouched, injected rows copied, fresh columns = Ldiag * Box–Muller(Philox(global
column, block, update))). Test infrastructure only: the product library has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

HC_DIR = os.path.join(ol.ROOT, "tests", "host_check")
CSRC = os.path.join(ol.ROOT, "assistedmanipulation_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def sc():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    so, src = os.path.join(HC_DIR, "libsample_check.so"), os.path.join(HC_DIR, "sample_check.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in ("sample_core.cuh", "host_math.h", "kernels.cuh", "rollout_core.cuh", "spatial.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-x", "c++", "-fPIC", "-shared", "-I" + CSRC, "-I" + CUDA_INC, "-o", so, src])
    lib = C.CDLL(so)
    lib.host_sample_quads.restype = C.c_longlong
    lib.host_sample_quads.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong,
                                      C.c_int, C.c_void_p, C.c_void_p]
    lib.host_sample_columns.restype = C.c_longlong
    lib.host_sample_columns.argtypes = lib.host_sample_quads.argtypes
    lib.host_quad_coordinates.argtypes = [C.c_longlong, C.c_longlong, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.host_philox.argtypes = [C.c_void_p] * 3
    lib.host_chase_coordinates.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


def test_philox_known_answers(sc):
    # Random123 kat_vectors, philox4x32 with 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        sc.host_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(x) for x in o) == want


def _philox_np(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint64 arrays holding 32-bit words (independent of the C++ under test)."""
    M0, M1, m = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xffffffff)
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & m, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & m
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m, (k1 + np.uint64(0xBB67AE85)) & m
    return c0, c1, c2, c3


def _box_muller_np(a, b):
    f = np.float32
    u1 = (a.astype(f) + f(0.5)) * f(2.3283064365386963e-10)
    u2 = (b.astype(f) + f(0.5)) * f(2.3283064365386963e-10)
    r = np.sqrt(f(-2.0) * np.log(u1))
    ang = f(6.283185307179586) * u2 - f(3.14159265358979)
    return r * np.cos(ang), r * np.sin(ang)


def _expected(T, k_begin, k_count, U, kept, ldiag, seed, upd, noise_source, injected, prefill):
    out = prefill.copy().reshape(k_count, T, 12)
    kg = np.arange(k_count) + k_begin
    col = (kg[:, None] * T + np.arange(T)[None, :]).astype(np.uint64)
    fresh = np.zeros((k_count, T, 12))
    if noise_source == 0:
        for b in range(3):
            r = _philox_np(col & np.uint64(0xffffffff), col >> np.uint64(32), np.full_like(col, b), np.full_like(col, upd & 0xffffffff), seed & 0xffffffff, seed >> 32)
            z0, z1 = _box_muller_np(r[0], r[1])
            z2, z3 = _box_muller_np(r[2], r[3])
            for i, z in enumerate((z0, z1, z2, z3)):
                fresh[:, :, 4 * b + i] = ldiag[4 * b + i] * z.astype(np.float64)   # (FP32 buffers: float(l) * z, one rounding apart)
    else:
        fresh = injected.reshape(k_count, T, 12).astype(np.float64)
    for kl in range(k_count):
        if kg[kl] == 0:
            out[kl] = 0.0
        elif kg[kl] == 1:
            out[kl] = -U.reshape(T, 12)
        elif not kept[kl]:
            out[kl] = fresh[kl]
    return out


@pytest.mark.parametrize("f32", [0, 1])
@pytest.mark.parametrize("k_begin,k_count,T", [(0, 37, 16), (0, 2, 8), (1, 5, 8), (19, 23, 10), (2 ** 33, 9, 64)])
@pytest.mark.parametrize("noise_source", [0, 1])
def test_quad_enumeration_matches_independent_construction(sc, f32, k_begin, k_count, T, noise_source):
    rng = np.random.default_rng(k_count * 131 + T)
    dt = np.float32 if f32 else np.float64
    U = rng.standard_normal(12 * T)
    kept = (rng.random(k_count) < 0.3).astype(np.uint8)
    ldiag = np.sqrt(np.array([0.1, 0.1, 0.2] + [7.5] * 7 + [0.0, 0.0]))
    seed, upd = 0x5EED00001234ABCD, 7
    inj_is_double = 1 if (noise_source == 1 and (k_count % 2 or not f32)) else 0     # host buffers are doubles, device buffers engine precision
    injected = rng.standard_normal(k_count * T * 12).astype(np.float64 if (inj_is_double or not f32) else np.float32)
    prefill = rng.standard_normal(k_count * T * 12).astype(dt)                     # what a kept rollout must keep
    out = prefill.copy()
    stored = sc.host_sample_quads(f32, inj_is_double, T, k_begin, k_count, U.ctypes.data, kept.ctypes.data, ldiag.ctypes.data, seed, upd, noise_source,
                                  injected.ctypes.data, out.ctypes.data)
    kg = np.arange(k_count) + k_begin
    rewritten = int(((kg < 2) | (kept == 0)).sum())
    assert stored == rewritten * T * 3
    want = _expected(T, k_begin, k_count, U, kept, ldiag, seed, upd, noise_source, injected, prefill.astype(np.float64)).astype(dt)
    got = out.reshape(k_count, T, 12)
    exact = (kg < 2) | (kept == 1) | (noise_source == 1)
    assert np.array_equal(got[exact], want[exact])
    # fresh columns: same counters and elements (libm vs numpy single-precision log / sin / cos differ by an ulp or two)
    assert np.allclose(got[~exact], want[~exact], rtol=2e-5, atol=2e-6)
    if noise_source == 0 and (~exact).any():
        assert np.all(got[~exact][..., 10:] == 0.0) and np.abs(got[~exact][..., :10]).min() > 0.0
    # the production kernel enumerates COLUMNS (one thread per NU values): same stores, bit for bit
    out2 = prefill.copy()
    stored2 = sc.host_sample_columns(f32, inj_is_double, T, k_begin, k_count, U.ctypes.data, kept.ctypes.data, ldiag.ctypes.data, seed, upd, noise_source,
                                     injected.ctypes.data, out2.ctypes.data)
    assert stored2 == stored and np.array_equal(out2, out)


def test_quad_coordinates_both_widths(sc):
    kl, t, b = C.c_longlong(), C.c_int(), C.c_int()
    rng = np.random.default_rng(1)
    for quads, T in [(3 * 64 * 4098, 64), (0x7fffffff, 128), (0x80000000, 64), (3 * 64 * 1048578 * 40, 64)]:
        for g in [0, 1, 2, 3, quads - 1, quads // 2] + [int(x) for x in rng.integers(0, quads, 200)]:
            sc.host_quad_coordinates(g, quads, T, C.byref(kl), C.byref(t), C.byref(b))
            assert (kl.value, t.value, b.value) == (g // (3 * T), (g // 3) % T, g % 3)
            assert (kl.value * T + t.value) * 12 + 4 * b.value == 4 * g     # the store address of the quad


def test_savitzky_golay_taps_known_answers(sc):
    # evaluated from gram_savitzky_golay.cpp:12-53
    sc.host_sg_weights.argtypes = [C.c_int, C.c_int, C.c_void_p]
    w = np.zeros(21)
    sc.host_sg_weights(10, 1, w.ctypes.data)
    assert np.allclose(w, 1.0 / 21.0, rtol=1e-14)
    w = np.zeros(5)
    sc.host_sg_weights(2, 2, w.ctypes.data)
    assert np.allclose(w, np.array([-3.0, 12.0, 17.0, 12.0, -3.0]) / 35.0, rtol=1e-14)
    for m, n in [(3, 2), (5, 3), (7, 4), (10, 2)]:     # the central tap row of the least-squares polynomial smoother
        w = np.zeros(2 * m + 1)
        sc.host_sg_weights(m, n, w.ctypes.data)
        x = np.arange(-m, m + 1, dtype=np.float64)
        A = np.vander(x, n + 1, increasing=True)
        assert np.allclose(w, np.linalg.pinv(A)[0], rtol=1e-10, atol=1e-13)


def test_noise_transform_is_the_reference_factor(sc):
    """V sqrt(Lambda), eigenvalues ascending (gaussian.hpp:48-55): L L^T = Sigma, column norms = sqrt of the sorted
    eigenvalues; for a diagonal Sigma the engine instead takes eps_i = sqrt(Sigma_ii) z_i (same distribution)."""
    sc.host_noise_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    sc.host_diagonal_noise_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    for n in (2, 5, 12):
        B = rng.standard_normal((n, n))
        S = B @ B.T + 0.1 * np.eye(n)
        Sc = np.asfortranarray(S)
        L = np.zeros((n, n), order="F")
        sc.host_noise_transform(n, Sc.ctypes.data, L.ctypes.data)
        assert np.allclose(L @ L.T, S, rtol=1e-11, atol=1e-12)
        ev = np.linalg.eigvalsh(S)
        assert np.allclose(np.linalg.norm(L, axis=0), np.sqrt(ev), rtol=1e-10)
        ld = np.zeros(n)
        assert sc.host_diagonal_noise_transform(n, Sc.ctypes.data, ld.ctypes.data) == 0
    d = np.array([0.1, 0.1, 0.2] + [7.5] * 7 + [0.0, 0.0])     # base.hpp:79-83
    S = np.asfortranarray(np.diag(d))
    L = np.zeros((12, 12), order="F")
    sc.host_noise_transform(12, S.ctypes.data, L.ctypes.data)
    assert np.allclose(L @ L.T, S, atol=1e-15)
    assert np.count_nonzero(L - np.diag(np.diag(L))) > 0           # the reference's factor of this matrix is a PERMUTED diagonal ...
    ld = np.zeros(12)
    assert sc.host_diagonal_noise_transform(12, S.ctypes.data, ld.ctypes.data) == 1
    assert np.array_equal(ld, np.sqrt(d))                          # ... the engine samples with the plain one
    assert np.array_equal(np.sort(np.linalg.norm(L, axis=0)), np.sort(ld))


@pytest.mark.parametrize("T", [1, 3, 4, 64, 130])
def test_chunk_major_enumeration_of_a_block_that_draws_its_own_noise(sc, T):
    """sample_core.cuh chase_sampler: sampling thread j of ns takes columns j, j + ns, ... of the block's chunk-major
    enumeration and adds one to the chunk's counter per column; the rollout warp starts a chunk when its counter reaches
    CHASE_COLUMNS. Every (rollout of the block, step) must be produced exactly once, every chunk must collect exactly
    CHASE_COLUMNS contributions (columns past the horizon count without being drawn), the chunks must complete in order of
    time, and the first chunk must be among the first columns handed out (one column time after the launch)."""
    steps, columns = sc.host_chase_steps(), sc.host_chase_columns()
    assert columns == 32 * steps
    chunks, ns = (T + steps - 1) // steps, 6 * 32
    seen, per_chunk, last_chunk = set(), [0] * chunks, -1
    c, r, t = C.c_int(), C.c_int(), C.c_int()
    for j in range(ns):
        for i in range(j, chunks * columns, ns):
            sc.host_chase_coordinates(i, C.byref(c), C.byref(r), C.byref(t))
            assert 0 <= r.value < 32 and c.value == t.value // steps == i // columns
            per_chunk[c.value] += 1
            if t.value < T:
                assert (r.value, t.value) not in seen
                seen.add((r.value, t.value))
    assert seen == {(r, t) for r in range(32) for t in range(T)}
    assert per_chunk == [columns] * chunks
    for i in range(chunks * columns):   # chunk-major: a chunk's columns are contiguous, so the first pass of the threads completes chunk 0
        sc.host_chase_coordinates(i, C.byref(c), C.byref(r), C.byref(t))
        assert c.value >= last_chunk
        last_chunk = c.value
    assert columns <= ns   # chunk 0 = the first `columns` indices, one per sampling thread
