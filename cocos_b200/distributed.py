"""One-process-per-GPU sharding of the Heston path (``torchrun`` / ``torch.distributed`` launch).

``torch.distributed`` is plumbing only: rendezvous, the broadcast of the NCCL unique
id and of the CUDA IPC handles, barriers and the max-over-ranks of timings.  The data
path has exactly one exchange -- the sum of the 2*nK payoff totals -- and it runs
INSIDE the simulation kernel: every rank's last block writes its totals into every
rank's exchange buffer over NVLink peer mappings and folds them in rank order
(``PeerExchange``, csrc/kernels.h).  Where peer mappings cannot be made, one
``ncclAllReduce`` on the kernel's stream does the same.  Every rank simulates a
disjoint, contiguous block of antithetic pairs from a disjoint Philox counter range, so
prices do not depend on the number of GPUs (up to fp64 summation order).  Stands in for
the process pool of cocos/multi_processing/device_pool.py:103-128.
"""
import ctypes
import math
import os
import typing as tp

import numpy as np

from . import _lib
from .multi_processing.utilities import shard_range


def env_rank_world() -> tp.Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def pair_shard(R: int, world_size: int, rank: int) -> tp.Tuple[int, int]:
    """(first, count) of the antithetic pairs rank ``rank`` simulates out of ceil(R/2)."""
    return shard_range((int(R) + 1) // 2, world_size, rank)


def broadcast_unique_id(group=None) -> bytes:
    """Rank 0 creates the NCCL unique id; everybody receives it through torch.distributed."""
    import torch.distributed as dist
    box = [None]
    if dist.get_rank(group) == 0:
        buf = ctypes.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        _lib.check(_lib.lib().cb_comm_unique_id(buf))
        box[0] = bytes(buf.raw)
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


class ShardedPricer:
    """Rank-local handle: context on this rank's GPU plus its NCCL communicator."""

    def __init__(self, device: int, rank: int, world_size: int, unique_id: tp.Optional[bytes] = None,
                 peer_exchange: bool = True, group=None):
        self.rank, self.world_size = rank, world_size
        self.ctx = _lib.context(device)
        self.comm = None
        self.peer_exchange = False
        self._group = group
        self._mapped = False          # exported or imported anything: the two-phase tear-down applies
        if world_size > 1:
            if unique_id is None or len(unique_id) != _lib.UNIQUE_ID_BYTES:
                raise ValueError("a 128-byte NCCL unique id is required when world_size > 1")
            handle = ctypes.c_void_p()
            _lib.check(_lib.lib().cb_comm_init_rank(self.ctx.handle, world_size, rank, unique_id,
                                                    ctypes.byref(handle)))
            self.comm = handle
            if peer_exchange and "COCOS_B200_NO_P2P" not in os.environ:
                self.peer_exchange = self._connect_peers(group)

    def _connect_peers(self, group) -> bool:
        """Exchange CUDA IPC handles of the ranks' exchange buffers; all ranks enable the fused exchange or none."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or self.world_size > 8:
            return False
        L = _lib.lib()
        buf = ctypes.create_string_buffer(_lib.P2P_HANDLE_BYTES)
        mine = bytes(buf.raw) if L.cb_comm_p2p_export(self.ctx.handle, self.comm, buf) == 0 else None
        self._mapped = mine is not None
        handles = [None] * self.world_size
        dist.all_gather_object(handles, mine, group=group)
        ok = all(h is not None for h in handles) and \
            L.cb_comm_p2p_import(self.ctx.handle, self.comm, b"".join(handles)) == 0
        votes = [None] * self.world_size
        dist.all_gather_object(votes, bool(ok), group=group)
        if not all(votes):
            # some rank could not map its peers: everybody unmaps what it did map NOW, and nobody frees an exported
            # buffer before all ranks have done so (the same two phases as close())
            L.cb_comm_p2p_disconnect(self.comm)
            dist.barrier(group=group)
            self._mapped = False
            return False
        _lib.check(L.cb_comm_p2p_enable(self.comm, 1))
        return True

    def launch(self, params: "_lib.HestonParams", R: int, nT: int, strikes, seed: int, subsequence: int,
               dtype_suffix: str = "f32", scheme: str = "euler", local_only: bool = False):
        """Asynchronous: this rank's shard of the simulation and the cross-rank sum, on one stream.
        ``local_only``: no exchange at all -- the result buffer keeps this rank's own shard sums."""
        k_arr, nK = _lib.strikes_array(strikes)
        first, count = pair_shard(R, self.world_size, self.rank)
        comm = None if local_only else self.comm
        if scheme == "conditional":
            _lib.check(_lib.lib().cb_heston_price_conditional_launch_f32(
                self.ctx.handle, ctypes.byref(params), int(R), int(nT), k_arr, nK, int(seed), int(subsequence), first,
                count, comm, None, None))
            return nK
        fn = getattr(_lib.lib(), f"cb_heston_price_launch_{dtype_suffix}")
        _lib.check(fn(self.ctx.handle, ctypes.byref(params), int(R), int(nT), k_arr, nK, int(seed), int(subsequence),
                      first, count, comm, None, None, None))
        return nK

    def pi_inside(self, n: int, seed: int, subsequence: int, antithetic: bool = True) -> int:
        """This rank's shard of the draws of a Monte-Carlo pi estimate plus the cross-rank sum of the inside count
        (fused into the kernel over the peer mappings, or one ncclAllReduce); every rank returns the total."""
        ndraw = (int(n) + 1) // 2 if antithetic else int(n)
        first, count = shard_range(ndraw, self.world_size, self.rank)
        L = _lib.lib()
        _lib.check(L.cb_pi_count_launch(self.ctx.handle, int(n), int(seed), int(subsequence), int(bool(antithetic)),
                                        first, count, self.comm))
        inside = ctypes.c_uint64(0)
        _lib.check(L.cb_pi_count_finish(self.ctx.handle, ctypes.byref(inside)))
        return inside.value

    def finish(self, nK: int):
        sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
        _lib.check(_lib.lib().cb_heston_price_finish(self.ctx.handle, nK, sums, sumsq))
        return np.array(sums), np.array(sumsq)

    def price(self, x0, v0, r, rho, sigma_v, kappa, v_bar, T, K, nT, R, seed=0, subsequence=0,
              dtype_suffix="f32") -> float:
        """simulate_and_compute_option_price for the WHOLE job; every rank returns the same price."""
        p = _lib.HestonParams(float(T), float(r), float(kappa), float(v_bar), float(sigma_v), float(rho),
                              float(x0), float(v0))
        nK = self.launch(p, R, nT, [K], seed, subsequence, dtype_suffix)
        sums, _ = self.finish(nK)
        return float(math.exp(-r * T) * sums[0] / int(R))

    def close(self):
        if self.comm is not None:
            if self.peer_exchange or self._mapped:
                # two phases: nobody frees its exported buffer while another rank still has it mapped
                _lib.lib().cb_comm_p2p_disconnect(self.comm)
                self.peer_exchange = False
                self._mapped = False
                try:
                    import torch.distributed as dist
                    if dist.is_available() and dist.is_initialized():
                        dist.barrier(group=self._group)
                except Exception:       # noqa: BLE001  (the process group may already be gone at interpreter exit)
                    pass
            _lib.lib().cb_comm_destroy(self.comm)
            self.comm = None
