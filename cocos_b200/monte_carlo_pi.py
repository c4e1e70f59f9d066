"""Monte-Carlo pi: examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py:18-33 as one fused kernel.

The reference draws two uniform arrays per batch, evaluates ``x*x + y*y <= 1`` and
means the mask (three array passes and two RNG launches per batch).  Here draws, test
and count happen in registers; ``antithetic=True`` gives the point layout of
``rand_antithetic([n], 0)`` (cocos/numerics/random.py:219-245) that BASELINE config 2 asks for.
"""
import ctypes
import math
import typing as tp

from . import _lib
from .device import current_context
from .numerics import random as _random


def count_inside(n: int, antithetic: bool = False, compute_device_pool=None) -> int:
    """Number of the n points that fall inside the quarter circle (sharded over a pool's GPUs when given)."""
    if compute_device_pool is not None:
        return compute_device_pool.pi_inside(int(n), _random.get_seed(), _random.next_subsequence(), antithetic)
    ctx = current_context()
    L = _lib.lib()
    ndraw = (n + 1) // 2 if antithetic else n
    _lib.check(L.cb_pi_count_launch(ctx.handle, int(n), _random.get_seed(), _random.next_subsequence(),
                                    int(bool(antithetic)), 0, ndraw, None))
    inside = ctypes.c_uint64(0)
    _lib.check(L.cb_pi_count_finish(ctx.handle, ctypes.byref(inside)))
    return inside.value


def estimate_pi(n: int, batches: int = 1, gpu: bool = True, antithetic: bool = False,
                compute_device_pool=None) -> float:
    """4 * (fraction of points inside the quarter circle), averaged over ``batches`` batches of ceil(n/batches).

    With ``compute_device_pool`` every batch is sharded over the pool's GPUs (disjoint draw ranges, one NCCL
    allreduce of the count) instead of the reference's map_reduce of whole batches
    (monte_carlo_pi_multi_gpu_example.py:185-188); the estimate does not depend on the number of GPUs."""
    if not gpu:
        raise NotImplementedError("cocos_b200 has no CPU branch (see oracle/ for the NumPy restatement)")
    n_per_batch = math.ceil(n / batches)
    pi = 0.0
    for _ in range(batches):
        pi += 4.0 * count_inside(n_per_batch, antithetic, compute_device_pool) / n_per_batch
    return pi / batches
