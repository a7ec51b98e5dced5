"""world_size-2 gloo test (CPU) of the host-side sharding protocol: two ranks each hold a contiguous
shard of the rollout costs and noise of one oracle update, exchange {-min,max} (MAX) and
{sum w, sum w*eps} (SUM) through torch.distributed, and must reproduce the oracle's gradient and
weights. Covers assistedmanipulation_b200/sharding.py, the logic bench.py and the engine use for N>1."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, T, nu, cost_scale, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from assistedmanipulation_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.load(out + "/oracle.npz")
    b, e = sharding.shard_range(K + 2, rank, world)
    costs, noise = data["costs"][b:e], data["noise"].reshape(K + 2, T * nu)[b:e]
    mm = sharding.all_reduce(dist, sharding.local_minmax(costs, rank, world), dist.ReduceOp.MAX)
    assert sharding.valid_total(mm) == 2.0
    w = sharding.local_weights(costs, mm, cost_scale)
    sums = sharding.all_reduce(dist, sharding.local_sums(w, noise), dist.ReduceOp.SUM)
    np.savez(out + "/rank%d.npz" % rank, minmax=mm, weights=w / sums[0], gradient=sums[1:] / sums[0], begin=b, end=e)
    dist.destroy_process_group()


@pytest.mark.parametrize("K", [61, 126])
def test_two_rank_exchange_reproduces_the_oracle(oracle, tmp_path, K):
    import torch.multiprocessing as mp
    import oracle_lib as ol
    from assistedmanipulation_b200 import abi, sharding
    T, nu = 20, 2
    holder = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.2, smoothing=None, control_bound=False)
    o = ol.Oracle(oracle, holder, abi.default_toy_objective())
    eps = np.random.default_rng(5).standard_normal((K + 2, T, nu))
    x0 = np.zeros(4)
    assert o.update(x0, 0.0, None, eps) == 0
    costs = o.read(abi.READ_COSTS, K + 2)
    noise = o.read(abi.READ_NOISE, (K + 2) * T * nu)
    weights, gradient, mm = o.read(abi.READ_WEIGHTS, K + 2), o.read(abi.READ_GRADIENT, nu * T), o.read(abi.READ_MINMAX, 2)
    o.close()
    # partition covers the index range exactly once, in order
    assert [sharding.shard_range(K + 2, r, 2) for r in range(2)] == [(0, (K + 2) // 2), ((K + 2) // 2, K + 2)]
    np.savez(str(tmp_path / "oracle.npz"), costs=costs, noise=noise)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, K, T, nu, 10.0, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(2)]
    for p in parts:
        assert -p["minmax"][0] == mm[0] and p["minmax"][1] == mm[1]   # min / max are exact
        assert np.allclose(p["gradient"], gradient, rtol=1e-12, atol=1e-15)
    assert np.allclose(np.concatenate([p["weights"] for p in parts]), weights, rtol=1e-12, atol=0)


def test_valid_rollouts_are_counted_after_the_exchange():
    """mppi.cpp:368-370 needs two valid rollouts in the WHOLE set: two ranks holding one each must not combine to one
    (saturating each rank's count before a MAX did that); the payload carries one slot per rank instead."""
    from assistedmanipulation_b200 import sharding
    nan = float("nan")
    parts = [sharding.local_minmax(np.array([nan, 3.0, nan]), 0, 2), sharding.local_minmax(np.array([nan, nan, 5.0]), 1, 2)]
    mm = np.max(np.stack(parts), axis=0)
    assert (-mm[0], mm[1]) == (3.0, 5.0) and sharding.valid_total(mm) == 2.0
    one = np.max(np.stack([sharding.local_minmax(np.array([nan, 3.0]), 0, 2), sharding.local_minmax(np.array([nan, nan]), 1, 2)]), axis=0)
    assert sharding.valid_total(one) == 1.0
    assert sharding.valid_total(sharding.local_minmax(np.array([1.0, 2.0, 3.0]))) == 2.0


def test_shard_ranges_are_contiguous_and_ordered():
    from assistedmanipulation_b200 import sharding
    for total in (4, 7, 4098, 1048578):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1


@pytest.mark.parametrize("world", [2, 3, 8])
def test_mailbox_protocol_under_adversarial_interleaving(world):
    sys.path.insert(0, ROOT)
    from assistedmanipulation_b200 import sharding
    """Host model of the device's peer-memory exchange (sharding.MailboxModel mirrors k_exchange step by step): under
    random schedules — including one rank racing as far ahead as the protocol lets it — no slot is overwritten before
    its owner combined it, every combine sees the payloads of its own exchange, and all ranks compute the same bits."""
    rng = np.random.default_rng(world)
    counts = [3, 9 + world, 6]
    updates = 12

    def payload(rank, attempt, kind):
        r = np.random.default_rng(1000 * attempt + 10 * kind + rank)
        return r.normal(size=counts[kind])

    for trial in range(30):
        m = sharding.MailboxModel(world, counts)
        m.payload_of = payload
        total = updates * len(counts)
        bias = rng.integers(0, world)          # one rank is scheduled far more often than the others
        guard = 0
        while min(m.exchange) < total:
            guard += 1
            assert guard < 200000, "deadlock in the model"
            runnable = [r for r in range(world) if m.exchange[r] < total]
            r = bias if (bias in runnable and rng.uniform() < 0.7) else runnable[rng.integers(0, len(runnable))]
            m.step(r)
        # a rank can never be two exchanges of the same (parity, kind) ahead of a peer: that is what makes two buffers enough
        for r in range(1, world):
            assert len(m.results[r]) == len(m.results[0]) == total
            for (a0, k0, v0), (a1, k1, v1) in zip(m.results[0], m.results[r]):
                assert (a0, k0) == (a1, k1) and np.array_equal(v0, v1)
        # and the combines are the reductions they stand for
        for attempt, kind, out in m.results[0]:
            parts = np.stack([payload(q, attempt, kind) for q in range(world)])
            want = parts.max(axis=0) if kind == sharding.EX_MINMAX else (parts.reshape(-1) if kind == sharding.EX_CAND else np.add.reduce(parts, axis=0))
            assert np.array_equal(out, want)


def test_mailbox_model_catches_a_single_buffered_mailbox():
    """The same model with ONE buffer per kind must trip its overwrite check: the double buffering is what the device relies on."""
    sys.path.insert(0, ROOT)
    from assistedmanipulation_b200 import sharding
    rng = np.random.default_rng(0)
    counts = [3]
    tripped = False
    for trial in range(50):
        m = sharding.MailboxModel(2, counts, parities=1)
        m.payload_of = lambda rank, attempt, kind: np.full(3, 10.0 * attempt + rank)
        try:
            for _ in range(400):
                r = 0 if rng.uniform() < 0.8 else 1     # rank 0 races ahead
                if m.exchange[r] < 8:
                    m.step(r)
        except AssertionError:
            tripped = True
            break
    assert tripped
