"""Host-side logic of a rollout set sharded over ranks (reference partition
src/controller/mppi.cpp:272-307 generalised from threads to GPUs).

Rank r of N owns the contiguous global sample indices [r*(K+2)//N, (r+1)*(K+2)//N) — contiguous so
that global index order (tie-breaking, rollouts 0 and 1 on rank 0) is preserved. One update needs
exactly two exchanges:

  1. MAX all-reduce of {-min_r, max_r, valid_r}   (both ends: the exponent is scaled by max - min,
     mppi.cpp:373,391-393, so a local-min rescale trick is not valid)
  2. SUM all-reduce of {sum_k w_k, sum_k w_k eps_k}  (1 + nu*T doubles)

after which every rank applies the same gradient step, smoothing and clamp redundantly.
The functions here are backend-agnostic (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(total, rank, world):
    """[begin, end) of the global sample indices rank owns; total = K + 2."""
    return total * rank // world, total * (rank + 1) // world


def local_minmax(costs, rank=0, world=1):
    """The engine's first exchange payload: {-min, max, 0, valid_0 .. valid_{world-1}} over the non-NaN local costs, with
    only this rank's valid slot filled (saturated at 2). Combined with an elementwise MAX, which gathers the disjoint slots;
    `valid_total` sums them — saturating per rank BEFORE a MAX would turn two ranks with one valid rollout each into one."""
    ok = ~np.isnan(costs)
    n = int(ok.sum())
    out = np.zeros(3 + world)
    out[0], out[1] = (-np.inf, -np.inf) if n == 0 else (-costs[ok].min(), costs[ok].max())
    out[3 + rank] = float(min(n, 2))
    return out


def valid_total(minmax):
    """valid rollouts of the whole set from a combined payload, saturated at 2 (mppi.cpp:368-370 needs two)"""
    return min(float(np.sum(minmax[3:])), 2.0)


def local_weights(costs, minmax, cost_scale):
    """Unnormalised likelihoods of the local shard given the GLOBAL {-min, max} (mppi.cpp:381-397)."""
    minimum, maximum = -minmax[0], minmax[1]
    difference = maximum - minimum
    w = np.exp(-cost_scale * (costs - minimum) / difference)
    w[np.isnan(costs)] = 0.0
    return w


def local_sums(weights, noise):
    """{sum w, sum_k w_k eps_k} for a shard; noise is [k_local, T*nu]."""
    return np.concatenate([[weights.sum()], weights @ noise])


def all_reduce(dist, array, op):
    """All-reduce a float64 numpy array through torch.distributed (gloo or NCCL)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64).copy())
    dist.all_reduce(t, op=op)
    return t.numpy()


# ---- host model of the peer-memory exchange (csrc/kernels.cuh: exchange_push / exchange_peer, k_misc.cu: k_exchange) ----
# (The two per-update exchanges carry their flag inside every 8-byte word of the payload — one "flag" per element instead
# of one per slot; the buffer-reuse argument below is the same: a word is overwritten only by the exchange two updates later.)
# The device protocol, one step at a time, so that its buffer-reuse argument can be checked under ANY interleaving of the
# ranks (tests/test_sharded_gloo.py): per (parity, kind) every rank's mailbox holds one slot and one flag per rank; a
# rank stores its payload into its slot of every peer's mailbox and then raises the flag with the sequence number of
# the exchange; it combines its own mailbox's slots in rank order once every flag carries that number.

EX_MINMAX, EX_SUMS, EX_CAND = 0, 1, 2


class MailboxModel:
    """All ranks' mailboxes plus one program counter per rank. `step(rank)` performs that rank's next atomic action:
    one store into one peer (payload, then flag) or — when every flag has arrived — the combine. A rank never blocks
    inside `step`; `step` returns False while the rank is waiting."""

    def __init__(self, world, counts, parities=2):
        self.world, self.counts, self.parities = world, counts, parities   # parities=1 models a single-buffered (broken) mailbox
        self.slots = [[[[None] * world for _ in counts] for _ in range(2)] for _ in range(world)]   # [owner][parity][kind][src]
        self.flags = [[[[0] * world for _ in counts] for _ in range(2)] for _ in range(world)]
        self.pc = [0] * world            # peers already served in the current exchange
        self.exchange = [0] * world      # exchanges completed per rank (attempt*len(kinds) + position)
        self.results = [[] for _ in range(world)]
        self.payload_of = None           # callable (rank, attempt, kind) -> np.ndarray

    def kinds(self):
        return list(range(len(self.counts)))

    def step(self, rank):
        kinds = self.kinds()
        attempt, kind = divmod(self.exchange[rank], len(kinds))
        parity, seq = attempt % self.parities, attempt * 4 + kind + 1
        payload = self.payload_of(rank, attempt, kind)
        peers = [p for p in range(self.world) if p != rank]
        if self.pc[rank] < len(peers):
            p = peers[self.pc[rank]]
            # the slot must not still be needed by its owner: the owner has combined every earlier exchange that used it
            pending = self.flags[p][parity][kind][rank]
            assert pending == 0 or self._consumed(p, pending), "rank %d overwrites a slot of rank %d that was not read yet" % (rank, p)
            self.slots[p][parity][kind][rank] = (seq, payload.copy())
            self.flags[p][parity][kind][rank] = seq
            self.pc[rank] += 1
            return True
        if any(self.flags[rank][parity][kind][q] != seq for q in peers):
            return False                 # spinning on a flag
        parts = []
        for q in range(self.world):
            if q == rank:
                parts.append(payload)
            else:
                got_seq, data = self.slots[rank][parity][kind][q]
                assert got_seq == seq, "rank %d read a payload of exchange %d while combining %d" % (rank, got_seq, seq)
                parts.append(data)
        parts = np.stack(parts)
        out = parts.max(axis=0) if kind == EX_MINMAX else (parts.reshape(-1) if kind == EX_CAND else np.add.reduce(parts, axis=0))
        self.results[rank].append((attempt, kind, out))
        self.exchange[rank] += 1
        self.pc[rank] = 0
        return True

    def _consumed(self, owner, seq):
        attempt, kind = (seq - 1) // 4, (seq - 1) % 4
        return self.exchange[owner] > attempt * len(self.counts) + kind
