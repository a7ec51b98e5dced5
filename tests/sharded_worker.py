"""Worker for tests/test_gpu_sharded.py::test_nccl_two_gpus (launched by torch.distributed.run):
each rank owns half of the rollouts on its own GPU; the library's NCCL all-reduces do the exchange.
Rank 0 compares against a single-GPU engine and the oracle."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import engine_lib as el  # noqa: E402
import oracle_lib as ol  # noqa: E402
from assistedmanipulation_b200 import abi  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K, T, nu = 2046, 32, 12
tp = abi.default_track_point()
e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, dynamics_mode=abi.DYNAMICS_FUSED, device=local, rank=rank, world_size=world, keep_best=20), tp)
exchange = os.environ.get("MPPI_B200_TEST_EXCHANGE", "nccl")   # "nccl" or "p2p" (NVLink peer-memory mailboxes)
assert abi.connect_ranks(e.lib, e.h, dist, torch, exchange) == 0, e.error()
x0 = abi.huddled_state()
Us, kept, best = [], [], []
for u in range(4):
    assert e.update(x0, 0.05 * u, None, seed=5) == 0, e.error()
    Us.append(e.read(abi.READ_OPTIMAL, nu * T))
    kept.append(e.read(abi.READ_KEPT, 20, np.int64))
    best.append(e.query(abi.QUERY_ARGMIN))
gathered = [None] * world
dist.all_gather_object(gathered, Us)
bests = [None] * world
dist.all_gather_object(bests, best)
if rank == 0:
    for other in gathered[1:]:
        for a, b in zip(Us, other):
            assert np.array_equal(a, b)
    whole = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, dynamics_mode=abi.DYNAMICS_FUSED, device=local, keep_best=20), tp)
    o = ol.Oracle(ol.load(), abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, threads=8, keep_best=20), tp)
    for u in range(4):
        assert whole.update(x0, 0.05 * u, None, seed=5) == 0
        Uw = whole.read(abi.READ_OPTIMAL, nu * T)
        assert np.array_equal(kept[u], whole.read(abi.READ_KEPT, 20, np.int64))   # the kept set is bit exact across shardings
        assert np.abs(Us[u] - Uw).max() <= 1e-12 * np.abs(Uw).max()
        assert all(b[u] == whole.query(abi.QUERY_ARGMIN) for b in bests), (u, bests)   # the best rollout is global on every rank
        assert o.update(x0, 0.05 * u, None, whole.read(abi.READ_NOISE, (K + 2) * T * nu)) == 0
        Uo = o.read(abi.READ_OPTIMAL, nu * T)
        assert np.abs(Us[u] - Uo).max() <= 1e-9 * np.abs(Uo).max()
    print("sharded ok", world, "ranks", exchange)
if exchange == "p2p":
    # a peer that does not arrive: the update returns an error after the 2 s time-out instead of hanging the device, and
    # the engine keeps the last good control sequence
    import time
    dist.barrier()
    if rank == 0:
        t0 = time.perf_counter()
        rc = e.update(x0, 0.05 * 4, None, seed=5)
        waited = time.perf_counter() - t0
        assert rc == abi.ERR_NCCL and "timed out" in e.error(), (rc, e.error())
        assert 1.0 < waited < 20.0, waited
        assert np.array_equal(e.read(abi.READ_OPTIMAL, nu * T), Us[-1])
        print("missing peer ok after %.1f s" % waited)
    dist.barrier()   # rank 1 kept its mailbox mapped meanwhile
e.close()
dist.destroy_process_group()
