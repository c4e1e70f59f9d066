"""``ComputeDevicePool``: map-reduce over the GPUs of one node.

Keeps the constructor, ``number_of_devices``, ``sync``, ``map_reduce`` and
``map_combine`` signatures of cocos/multi_processing/device_pool.py:58-341, but not
its machinery: the reference binds one OS process to each GPU (loky/pathos,
device_pool.py:103-128) and ships pickled closures through pipes.  Here one process
drives every GPU:

* generic ``f``: batch i runs on device i mod G from one host thread per device (the
  current device is thread-local), results are folded on the host from the left in
  ``args_list`` order (device_pool.py:246-251), transfer hooks and the sync()s around
  ``f`` are honoured (device_pool.py:214-222);
* the Heston pricing path (``heston_sums``): every GPU simulates a disjoint block of
  antithetic pairs from a disjoint Philox counter range on its own stream and ONE
  ncclAllReduce combines the payoff sums and sums of squares -- no host fold, no
  pickling, no inter-GPU traffic during the simulation.
"""
import ctypes
import threading
import typing as tp
from concurrent.futures import ThreadPoolExecutor

from .. import _lib
from ..device import ComputeDevice, ComputeDeviceManager
from ..numerics import random as _random
from .utilities import (MultiprocessingPoolType, ParameterTransferFunction, ResultType,
                        _extract_arguments_and_number_of_batches, ordered_left_fold, shard_range)

ComputeDeviceFilter = tp.Callable[[ComputeDevice], bool]


def exclude_intel_devices(compute_device: ComputeDevice) -> bool:
    """Kept for signature compatibility (device_pool.py:47-55); a B200 node has no Intel GPU."""
    return 'intel' not in compute_device.name.lower()


def _get_set_of_compute_devices_from_iterable(
        compute_devices: tp.Iterable[tp.Union[int, ComputeDevice]]) -> tp.FrozenSet[ComputeDevice]:
    result = []
    for i, compute_device in enumerate(compute_devices):
        if isinstance(compute_device, int):
            compute_device = ComputeDeviceManager.get_compute_device(compute_device)
        elif not isinstance(compute_device, ComputeDevice):
            raise TypeError(f"Every element in compute_devices must be of type ComputeDevice or of type int. "
                            f"Entry {i} is of type {type(compute_device)}.")
        result.append(compute_device)
    return frozenset(result)


class ComputeDevicePool:
    def __init__(self,
                 compute_devices: tp.Optional[tp.Iterable[tp.Union[int, ComputeDevice]]] = None,
                 compute_device_filter: tp.Optional[ComputeDeviceFilter] = exclude_intel_devices,
                 multiprocessing_pool_type: MultiprocessingPoolType = MultiprocessingPoolType.default()) -> None:
        if compute_devices is None:
            compute_devices = ComputeDeviceManager.get_compute_devices()
        devices = _get_set_of_compute_devices_from_iterable(compute_devices)
        if compute_device_filter is not None:
            devices = frozenset(filter(compute_device_filter, devices))
        if not devices:
            raise ValueError("the compute device pool is empty")
        self._compute_devices = devices
        self._ordered = tuple(sorted(devices, key=lambda d: d.id))
        self._multiprocessing_pool_type = multiprocessing_pool_type
        self._contexts = [_lib.context(d.id) for d in self._ordered]
        self._comms: tp.Optional[tp.List[ctypes.c_void_p]] = None
        self._threads: tp.Optional[ThreadPoolExecutor] = None

    # ---- reference surface ------------------------------------------------
    @property
    def compute_devices(self) -> tp.FrozenSet[ComputeDevice]:
        return self._compute_devices

    @property
    def number_of_devices(self) -> int:
        return len(self._compute_devices)

    @property
    def multiprocessing_pool_type(self) -> MultiprocessingPoolType:
        return self._multiprocessing_pool_type

    def sync(self):
        for ctx in self._contexts:
            ctx.sync()

    def _run_batches(self, f, h2d, d2h, args_list, kwargs_list):
        parent_seed = _random.get_seed()
        epoch = _random.next_subsequence()      # a fan-out consumes one launch slot of the parent stream
        devices = self._ordered

        def synced_f(index, args, kwargs):
            device = devices[index % len(devices)]
            ComputeDeviceManager.set_compute_device(device.id)
            # every batch draws from its own stream, whatever thread or order it runs in
            _random.set_state(_random.derive_seed(parent_seed, epoch, index))
            if h2d is not None:
                args, kwargs = h2d(*args, **kwargs)
            device.sync()
            result = f(*args, **kwargs)
            if d2h is not None:
                result = d2h(result)
            device.sync()
            return index, result

        if self._threads is None:
            self._threads = ThreadPoolExecutor(max_workers=len(devices), thread_name_prefix="cocos-b200-gpu")
        # batch i is pinned to device i mod G: one in-flight batch per device at a time
        lanes = [[] for _ in devices]
        for i, (a, k) in enumerate(zip(args_list, kwargs_list)):
            lanes[i % len(devices)].append((i, tuple(a), dict(k)))
        futures = [self._threads.submit(lambda lane=lane: [synced_f(*job) for job in lane]) for lane in lanes]
        results = []
        for fut in futures:
            results.extend(fut.result())
        return results

    def map_reduce(self,
                   f: tp.Callable[..., ResultType],
                   reduction: tp.Callable[[ResultType, ResultType], ResultType],
                   initial_value: ResultType,
                   host_to_device_transfer_function: tp.Optional[ParameterTransferFunction] = None,
                   device_to_host_transfer_function: tp.Optional[tp.Callable[[ResultType], ResultType]] = None,
                   args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
                   kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
                   number_of_batches: tp.Optional[int] = None) -> ResultType:
        """Evaluate ``f`` per batch across the pool and fold the results from the left in list order."""
        args_list, kwargs_list, number_of_batches = _extract_arguments_and_number_of_batches(
            args_list=args_list, kwargs_list=kwargs_list, number_of_batches=number_of_batches)
        results = self._run_batches(f, host_to_device_transfer_function, device_to_host_transfer_function,
                                    args_list, kwargs_list)
        return ordered_left_fold(results, reduction, initial_value)

    def map_combine(self,
                    f: tp.Callable[..., ResultType],
                    combination: tp.Callable[[tp.Iterable[ResultType]], ResultType],
                    host_to_device_transfer_function: tp.Optional[ParameterTransferFunction] = None,
                    device_to_host_transfer_function: tp.Optional[tp.Callable[[ResultType], ResultType]] = None,
                    args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
                    kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
                    number_of_batches: tp.Optional[int] = None) -> ResultType:
        """Evaluate ``f`` per batch and hand the ordered result list to ``combination`` (device_pool.py:255-341)."""
        args_list, kwargs_list, number_of_batches = _extract_arguments_and_number_of_batches(
            args_list=args_list, kwargs_list=kwargs_list, number_of_batches=number_of_batches)
        results = self._run_batches(f, host_to_device_transfer_function, device_to_host_transfer_function,
                                    args_list, kwargs_list)
        return combination([value for _, value in sorted(results, key=lambda item: item[0])])

    # ---- fused fast path ----------------------------------------------------
    def _communicators(self):
        if self._comms is None and self.number_of_devices > 1:
            n = self.number_of_devices
            ctxs = (ctypes.c_void_p * n)(*[c.handle for c in self._contexts])
            comms = (ctypes.c_void_p * n)()
            _lib.check(_lib.lib().cb_comm_init_all(n, ctxs, comms))
            self._comms = [ctypes.c_void_p(c) for c in comms]
        return self._comms

    @property
    def peer_exchange_active(self) -> bool:
        """True when the cross-GPU sum of the fused launches runs inside the kernel over NVLink peer mappings
        (every device pair has peer access); False: one ncclAllReduce follows each kernel."""
        comms = self._communicators()
        if comms is None:
            return False
        active = ctypes.c_int(0)
        _lib.check(_lib.lib().cb_comm_p2p_active(comms[0], ctypes.byref(active)))
        return bool(active.value)

    def heston_sums(self, params: "_lib.HestonParams", R: int, N: int, strikes, seed: int, subsequence: int,
                    dtype_suffix: str = "f32", scheme: str = "euler",
                    spec: tp.Optional["_lib.PayoffSpec"] = None) -> tp.Tuple[tp.List[float], tp.List[float]]:
        """Shard ceil(R/2) antithetic pairs over the pool; one NCCL allreduce; returns (sums, sumsq).
        ``spec`` selects a path-dependent (or put) payoff of the Euler kernel."""
        L = _lib.lib()
        k_arr, nK = _lib.strikes_array(strikes)
        sums = (ctypes.c_double * nK)()
        sumsq = (ctypes.c_double * nK)()
        G = self.number_of_devices
        comms = self._communicators()
        if scheme == "euler":
            # the whole job in one C call: shards, launches, grouped allreduce, one host read
            ctxs, comm_arr = self._handle_arrays(comms)
            multi = getattr(L, f"cb_heston_price_multi_{dtype_suffix}")
            _lib.check(multi(G, ctxs, comm_arr, ctypes.byref(params), R, N, k_arr, nK, seed, subsequence,
                             ctypes.byref(spec) if spec is not None else None, sums, sumsq))
            return list(sums), list(sumsq)
        if scheme != "conditional":
            raise ValueError(f"unknown scheme {scheme!r}")
        if spec is not None or dtype_suffix != "f32":
            raise TypeError("the conditional scheme prices European calls in float32 only")
        pairs = (R + 1) // 2
        if comms is not None:
            _lib.check(L.cb_group_start())
        try:
            for g, ctx in enumerate(self._contexts):
                first, count = shard_range(pairs, G, g)
                _lib.check(L.cb_heston_price_conditional_launch_f32(
                    ctx.handle, ctypes.byref(params), R, N, k_arr, nK, seed, subsequence, first, count,
                    comms[g] if comms is not None else None, None, None))
        finally:
            if comms is not None:
                _lib.check(L.cb_group_end())
        # every rank holds the reduced sums: read them from the first device, just drain the others
        _lib.check(L.cb_heston_price_finish(self._contexts[0].handle, nK, sums, sumsq))
        for ctx in self._contexts[1:]:
            ctx.sync()
        return list(sums), list(sumsq)

    def _handle_arrays(self, comms):
        n = self.number_of_devices
        ctxs = (ctypes.c_void_p * n)(*[c.handle for c in self._contexts])
        comm_arr = (ctypes.c_void_p * n)(*[c.value for c in comms]) if comms is not None else None
        return ctxs, comm_arr

    def pi_inside(self, n: int, seed: int, subsequence: int, antithetic: bool) -> int:
        """Shard the draws of a Monte-Carlo pi estimate over the pool; one NCCL allreduce of the inside count."""
        ctxs, comm_arr = self._handle_arrays(self._communicators())
        inside = ctypes.c_uint64(0)
        _lib.check(_lib.lib().cb_pi_count_multi(self.number_of_devices, ctxs, comm_arr, int(n), seed, subsequence,
                                                int(bool(antithetic)), ctypes.byref(inside)))
        return inside.value

    def close(self):
        if self._comms is not None:
            for c in self._comms:
                _lib.lib().cb_comm_destroy(c)
            self._comms = None
        if self._threads is not None:
            self._threads.shutdown(wait=True)
            self._threads = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
