"""map_reduce / map_combine over CPU cores (cocos/multi_processing/multi_core_batch_processing.py:10-175).

Host-side helper kept for API completeness; it never touches the GPU path.  loky and
pathos are not required: a spawn-context ProcessPoolExecutor with cloudpickle
serialisation (lambdas and closures work, as they do under loky) is used instead.
"""
import multiprocessing
import typing as tp
from concurrent.futures import ProcessPoolExecutor

from .utilities import (MultiprocessingPoolType, ResultType, _extract_arguments_and_number_of_batches,
                        ordered_left_fold)

_executor = None


def _get_executor() -> ProcessPoolExecutor:
    global _executor
    if _executor is None:
        _executor = ProcessPoolExecutor(max_workers=multiprocessing.cpu_count(),
                                        mp_context=multiprocessing.get_context("spawn"))
    return _executor


def _run_pickled(payload: bytes):
    import cloudpickle
    index, f, args, kwargs = cloudpickle.loads(payload)
    return index, f(*args, **kwargs)


def _submit_all(f, args_list, kwargs_list):
    import cloudpickle
    ex = _get_executor()
    futures = [ex.submit(_run_pickled, cloudpickle.dumps((i, f, tuple(a), dict(k))))
               for i, (a, k) in enumerate(zip(args_list, kwargs_list))]
    return [fut.result() for fut in futures]


def map_reduce_multicore(
        f: tp.Callable[..., ResultType],
        reduction: tp.Callable[[ResultType, ResultType], ResultType],
        initial_value: ResultType,
        args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
        kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
        number_of_batches: tp.Optional[int] = None,
        multiprocessing_pool_type: MultiprocessingPoolType = MultiprocessingPoolType.FUTURES) -> ResultType:
    args_list, kwargs_list, number_of_batches = _extract_arguments_and_number_of_batches(
        args_list=args_list, kwargs_list=kwargs_list, number_of_batches=number_of_batches)
    return ordered_left_fold(_submit_all(f, args_list, kwargs_list), reduction, initial_value)


def map_combine_multicore(
        f: tp.Callable[..., ResultType],
        combination: tp.Callable[[tp.Iterable[ResultType]], ResultType],
        args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
        kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
        number_of_batches: tp.Optional[int] = None,
        multiprocessing_pool_type: MultiprocessingPoolType = MultiprocessingPoolType.FUTURES) -> ResultType:
    args_list, kwargs_list, number_of_batches = _extract_arguments_and_number_of_batches(
        args_list=args_list, kwargs_list=kwargs_list, number_of_batches=number_of_batches)
    results = sorted(_submit_all(f, args_list, kwargs_list), key=lambda item: item[0])
    return combination([value for _, value in results])
