"""Sharded rollout sets on the GPU.

* test_two_shards_on_one_gpu: two engines (rank 0 / rank 1 of world 2) on ONE device driven through
  the split C ABI (update_begin / update_weights / update_finish) with the two exchanges done on the
  host — the ranks run as sequential phases, never as kernels waiting on each other.
* test_nccl_two_gpus: the in-library NCCL exchange, one process per GPU (needs >= 2 GPUs).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu


def _cudart():
    import torch  # noqa: F401  (loads the CUDA runtime the wheels ship)
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    import glob
    import torch
    for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        return C.CDLL(p)
    raise RuntimeError("libcudart not found")


def _exchange(rt, engines, which, op):
    ptrs, counts = [], []
    for e in engines:
        mm, sums, n1, n2 = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        assert e.lib.mppi_b200_reduce_buffers(e.h, C.byref(mm), C.byref(n1), C.byref(sums), C.byref(n2)) == 0
        ptrs.append(mm if which == 0 else sums)
        counts.append(n1.value if which == 0 else n2.value)
        assert e.lib.mppi_b200_synchronize(e.h) == 0
    bufs = []
    for p, n in zip(ptrs, counts):
        b = np.zeros(n)
        assert rt.cudaMemcpy(b.ctypes.data_as(C.c_void_p), p, C.c_size_t(n * 8), 2) == 0
        bufs.append(b)
    red = op(np.stack(bufs), axis=0)
    for p, n in zip(ptrs, counts):
        assert rt.cudaMemcpy(p, red.ctypes.data_as(C.c_void_p), C.c_size_t(n * 8), 1) == 0
    return red


@pytest.mark.parametrize("source", ["philox", "host"])
def test_two_shards_on_one_gpu(oracle, source):
    import engine_lib as el
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    K, T, nu = 510, 32, 12
    tp = abi.default_track_point()
    mk = lambda r, w: abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, dynamics_mode=abi.DYNAMICS_FUSED, rank=r, world_size=w)
    shards = [el.Engine(mk(0, 2), tp), el.Engine(mk(1, 2), tp)]
    whole = el.Engine(mk(0, 1), tp)
    o = ol.Oracle(oracle, abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, threads=8), tp)
    assert [(e.query(abi.QUERY_LOCAL_BEGIN), e.query(abi.QUERY_LOCAL_COUNT)) for e in shards] == [(0, 256), (256, 256)]
    x0 = abi.huddled_state()
    rng = np.random.default_rng(2)
    for u in range(3):
        t = 0.05 * u
        eps = rng.standard_normal((K + 2, T, nu)) * np.sqrt(abi.FRANKA_COVARIANCE_DIAG)
        st = np.ascontiguousarray(x0)
        for e in shards:
            if source == "host":
                rc = e.lib.mppi_b200_update_begin(e.h, el.ptr(st), t, None, eps.ctypes.data_as(C.c_void_p), abi.NOISE_HOST, 0)
            else:
                rc = e.lib.mppi_b200_update_begin(e.h, el.ptr(st), t, None, None, abi.NOISE_PHILOX, 77)
            assert rc == 0, e.error()
        _exchange(rt, shards, 0, np.max)
        for e in shards:
            assert e.lib.mppi_b200_update_weights(e.h) == 0
        _exchange(rt, shards, 1, np.sum)
        for e in shards:
            assert e.lib.mppi_b200_update_finish(e.h) == 0, e.error()
        assert whole.update(x0, t, None, eps if source == "host" else None, seed=77) == 0
        Uw = whole.read(abi.READ_OPTIMAL, nu * T)
        U0, U1 = shards[0].read(abi.READ_OPTIMAL, nu * T), shards[1].read(abi.READ_OPTIMAL, nu * T)
        assert np.array_equal(U0, U1)                                   # every rank applies the same update
        assert np.abs(U0 - Uw).max() <= 1e-12 * np.abs(Uw).max()         # only the summation order differs
        cw = whole.read(abi.READ_COSTS, K + 2)
        cs = np.concatenate([e.read(abi.READ_COSTS, 256) for e in shards])
        assert np.allclose(cs, cw, rtol=1e-12, atol=0)                   # U_shift differs by rounding only
        # the best rollout is global: every rank reports the lowest index of the minimum over ALL shards
        assert shards[0].query(abi.QUERY_ARGMIN) == shards[1].query(abi.QUERY_ARGMIN) == int(np.argmin(cs))
        nw = whole.read(abi.READ_NOISE, (K + 2) * T * nu).reshape(K + 2, -1)
        ns = np.concatenate([e.read(abi.READ_NOISE, 256 * T * nu).reshape(256, -1) for e in shards])
        assert np.array_equal(ns[2:], nw[2:])                            # the Philox stream does not depend on the sharding
        noise = nw.reshape(-1) if source == "philox" else eps
        assert o.update(x0, t, None, noise) == 0
        Uo = o.read(abi.READ_OPTIMAL, nu * T)
        assert np.abs(U0 - Uo).max() <= 1e-9 * np.abs(Uo).max()
    for e in shards + [whole]:
        e.close()
    o.close()


def test_sharded_keep_best_needs_the_library_exchange():
    # the warm start over a sharded set all-gathers candidates through the library's communicator
    import engine_lib as el
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 64, 0.2, keep_best=4, rank=0, world_size=2)
    e = el.Engine(h, abi.default_toy_objective())
    st = np.zeros(4)
    assert e.lib.mppi_b200_update_begin(e.h, el.ptr(st), 0.0, None, None, abi.NOISE_PHILOX, 0) == abi.ERR_UNSUPPORTED
    e.close()


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_two_gpus_in_library_exchange(exchange):
    """One process per GPU; the library exchanges min/max, weighted sums and warm-start candidates itself — through NCCL
    or through its own kernels over NVLink peer memory (IPC-mapped mailboxes). Rank 0 compares with one GPU and the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(ol.ROOT, "tests", "sharded_worker.py")
    env = dict(os.environ, MPPI_B200_TEST_EXCHANGE=exchange)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611" if exchange == "nccl" else "29612", worker], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded ok 2 ranks " + exchange in out.stdout
    if exchange == "p2p":
        assert "missing peer ok" in out.stdout
