"""Multi-GPU plumbing: chains shard trivially across ranks (independent Markov processes, Philox keyed by the
global chain id), so the only cross-rank step is the final reduction of moment sums.  Uses
torch.distributed when it is initialised (NCCL on GPUs, gloo in the CPU tests); a no-op in a single process."""
from __future__ import annotations

import numpy as np


def shard(n_chains_total: int, rank: int, world: int):
    """Contiguous block of chains owned by `rank`: returns (chain_offset, n_chains)."""
    base, rem = divmod(int(n_chains_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def moment_sums(mean, second):
    """Per-rank sufficient statistics of the chain time-averages: rows (sum_c m_c, sum_c m_c^2, sum_c s_c, count)
    with m_c = time average of x, s_c = time average of x^2 of chain c.  Works on numpy arrays or torch tensors."""
    n = mean.shape[0]
    if hasattr(mean, "new_full"):
        import torch
        return torch.stack([mean.sum(0), (mean * mean).sum(0), second.sum(0), mean.new_full((mean.shape[1],), float(n))])
    return np.stack([mean.sum(0), (mean * mean).sum(0), second.sum(0), np.full(mean.shape[1], float(n))])


def all_reduce_sums(sums):
    """Sum the moment sums over all ranks (the only collective of the whole path).  `sums` is a torch tensor on the
    rank's device (NCCL) or CPU (gloo); returned unchanged when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
    except ImportError:
        return sums
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums)
    return sums


def ess_from_sums(sums):
    """Pooled mean / variance and cross-chain ESS per coordinate from the (reduced) moment sums:
    ESS_i = C * Var_pi(x_i) / Var_c(m_{c,i})."""
    s = np.asarray(sums.cpu() if hasattr(sums, "cpu") else sums, dtype=np.float64)
    C = s[3]
    mbar = s[0] / C
    var_between = (s[1] / C - mbar**2) * C / np.maximum(C - 1.0, 1.0)
    pooled_var = s[2] / C - mbar**2
    return {"mean": mbar, "var": pooled_var, "ess": C * pooled_var / var_between, "chains": C[0]}
