"""Multi-GPU plumbing: chains shard trivially across ranks (independent Markov processes, Philox keyed by the
global chain id), so the only cross-rank step is the final reduction of moment sums.

Both steps of that reduction live in the C ABI (`pdmpflux_moments_reduce`: one fused kernel for the 4 x d sufficient
statistics; `pdmpflux_moments_allreduce`: NCCL, communicator owned by the library), so a Julia host calls exactly what
this module calls.  The host's only job is to carry the 128-byte NCCL unique id from rank 0 to the others; here that is
a torch.distributed broadcast (any transport would do).  The numpy / CPU-tensor branches below exist for the CPU test
suite (gloo, no GPU); they are not a fallback of the product path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def shard(n_chains_total: int, rank: int, world: int):
    """Contiguous block of chains owned by `rank`: returns (chain_offset, n_chains)."""
    base, rem = divmod(int(n_chains_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def moment_sums(mean, second):
    """Per-rank sufficient statistics of the chain time-averages: rows (sum_c m_c, sum_c m_c^2, sum_c s_c, count)
    with m_c = time average of x, s_c = time average of x^2 of chain c.  numpy arrays or CPU tensors (host-side
    bookkeeping and the CPU tests); device data goes through `moment_sums_device`."""
    n = mean.shape[0]
    if hasattr(mean, "new_full"):
        import torch
        return torch.stack([mean.sum(0), (mean * mean).sum(0), second.sum(0), mean.new_full((mean.shape[1],), float(n))])
    return np.stack([mean.sum(0), (mean * mean).sum(0), second.sum(0), np.full(mean.shape[1], float(n))])


def moment_sums_device(m1, m2, T, out, stream=None):
    """`pdmpflux_moments_reduce` on device buffers (torch CUDA tensors or raw pointers): m1, m2 [C][d] time integrals,
    T [C] their time spans (None: already averages), out [4][d].  One fused kernel pair, fixed summation order."""
    ptr = lambda a: None if a is None else (a.data_ptr() if hasattr(a, "data_ptr") else int(a))
    n_chains, d = m1.shape
    _lib.check(_lib.lib().pdmpflux_moments_reduce(int(d), int(n_chains), ptr(m1), ptr(m2), ptr(T), ptr(out), 1, stream))
    return out


class Comm:
    """The library-owned NCCL communicator of this rank (`pdmpflux_comm_*`).  World size / rank come from
    torch.distributed when it is initialised (the unique id travels through its broadcast), else a single rank."""

    def __init__(self, device=None):
        self.rank, self.world, self._h = 0, 1, C.c_void_p()
        try:
            import torch
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            dist = None
        lib = _lib.lib()
        idbuf = (C.c_ubyte * 128)()
        if self.world > 1:
            if self.rank == 0:
                _lib.check(lib.pdmpflux_comm_unique_id(idbuf, 128))
            t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
            dist.broadcast(t, src=0)
            raw = bytes(t.cpu().tolist())
            idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
        _lib.check(lib.pdmpflux_comm_create(idbuf, self.world, self.rank, C.byref(self._h)))

    def all_reduce(self, sums, stream=None):
        """In-place sum over ranks of a float64 CUDA tensor (`pdmpflux_moments_allreduce`, ncclAllReduce)."""
        _lib.check(_lib.lib().pdmpflux_moments_allreduce(self._h, sums.data_ptr(), sums.numel(), stream))
        return sums

    def close(self):
        if self._h:
            _lib.lib().pdmpflux_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def all_reduce_sums(sums, comm=None, stream=None):
    """Sum the moment sums over all ranks (the only collective of the whole path).  CUDA tensors go through the
    library's own NCCL communicator (`comm`, a `Comm`); CPU tensors (the gloo tests) through torch.distributed;
    unchanged in a single process."""
    if comm is not None and getattr(sums, "is_cuda", False):
        return comm.all_reduce(sums, stream)
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
    C_ = s[3]
    mbar = s[0] / C_
    var_between = (s[1] / C_ - mbar**2) * C_ / np.maximum(C_ - 1.0, 1.0)
    pooled_var = s[2] / C_ - mbar**2
    return {"mean": mbar, "var": pooled_var, "ess": C_ * pooled_var / var_between, "chains": C_[0]}
