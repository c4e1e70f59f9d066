"""Sequential map_reduce / map_combine on the current device
(cocos/multi_processing/single_device_batch_processing.py:12-154)."""
import typing as tp

from ..device import sync
from .utilities import ResultType, ParameterTransferFunction, _extract_arguments_and_number_of_batches


def _run(f, h2d, d2h, args, kwargs):
    if h2d is not None:
        args, kwargs = h2d(*args, **kwargs)
    sync()
    result = f(*args, **kwargs)
    if d2h is not None:
        result = d2h(result)
    sync()
    return result


def map_reduce_single_device(
        f: tp.Callable[..., ResultType],
        reduction: tp.Callable[[ResultType, ResultType], ResultType],
        initial_value: ResultType,
        host_to_device_transfer_function: tp.Optional[ParameterTransferFunction] = None,
        device_to_host_transfer_function: tp.Optional[tp.Callable[[ResultType], ResultType]] = None,
        args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
        kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
        number_of_batches: tp.Optional[int] = None) -> ResultType:
    args_list, kwargs_list, number_of_batches = _extract_arguments_and_number_of_batches(
        args_list=args_list, kwargs_list=kwargs_list, number_of_batches=number_of_batches)
    result = initial_value
    for args, kwargs in zip(args_list, kwargs_list):
        result = reduction(result, _run(f, host_to_device_transfer_function, device_to_host_transfer_function,
                                        args, kwargs))
    return result


def map_combine_single_device(
        f: tp.Callable[..., ResultType],
        combination: tp.Callable[[tp.Iterable[ResultType]], ResultType],
        host_to_device_transfer_function: tp.Optional[ParameterTransferFunction] = None,
        device_to_host_transfer_function: tp.Optional[tp.Callable[[ResultType], ResultType]] = None,
        args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
        kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
        number_of_batches: tp.Optional[int] = None) -> ResultType:
    args_list, kwargs_list, number_of_batches = _extract_arguments_and_number_of_batches(
        args_list=args_list, kwargs_list=kwargs_list, number_of_batches=number_of_batches)
    return combination([_run(f, host_to_device_transfer_function, device_to_host_transfer_function, a, k)
                        for a, k in zip(args_list, kwargs_list)])
