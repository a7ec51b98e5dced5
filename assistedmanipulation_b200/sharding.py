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


def local_minmax(costs):
    """{-min, max, valid} over the non-NaN local costs (valid saturates at 2, like the engine)."""
    ok = ~np.isnan(costs)
    n = int(ok.sum())
    if n == 0:
        return np.array([-np.inf, -np.inf, 0.0])
    return np.array([-costs[ok].min(), costs[ok].max(), float(min(n, 2))])


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
