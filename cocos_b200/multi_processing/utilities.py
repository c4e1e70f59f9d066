"""cocos/multi_processing/utilities.py: slice generators and argument normalisation."""
import math
import typing as tp
from enum import Enum


class MultiprocessingPoolType(Enum):
    """The reference offers loky or pathos process pools (utilities.py:23-40).  Here the
    device pool is in-process (streams + NCCL) and the CPU pool is concurrent.futures."""
    LOKY = 1
    PATHOS = 2
    STREAMS = 3
    FUTURES = 4

    @staticmethod
    def default() -> 'MultiprocessingPoolType':
        return MultiprocessingPoolType.STREAMS


def generate_slices_with_batch_size(n: int, batch_size: int) -> tp.Tuple[tp.Tuple[int, int], ...]:
    """Partition range(n) into consecutive (begin, end) ranges of length <= batch_size."""
    if batch_size <= 0:
        raise ValueError("batch_size must be positive")
    return tuple((lo, min(lo + batch_size, n)) for lo in range(0, n, batch_size))


def generate_slices_with_number_of_batches(n: int, number_of_batches: int) -> tp.Tuple[tp.Tuple[int, int], ...]:
    """Partition range(n) into ranges of length ceil(n/number_of_batches) (the last may be shorter)."""
    return generate_slices_with_batch_size(n=n, batch_size=math.ceil(n / number_of_batches))


ResultType = tp.TypeVar('ResultType')
ParameterTransferFunction = tp.Callable[..., tp.Tuple[tp.Sequence, tp.Dict[str, tp.Any]]]


def _extract_arguments_and_number_of_batches(
        args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
        kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
        number_of_batches: tp.Optional[int] = None):
    """utilities.py:88-106: derive the batch count and fill the missing argument lists."""
    if number_of_batches is None:
        if args_list is not None:
            number_of_batches = len(args_list)
        elif kwargs_list is not None:
            number_of_batches = len(kwargs_list)
        else:
            raise ValueError('Number_of_batches must be defined if both args_list and kwargs_list are empty')
    if args_list is None:
        args_list = number_of_batches * [list()]
    if kwargs_list is None:
        kwargs_list = number_of_batches * [dict()]
    return args_list, kwargs_list, number_of_batches


def ordered_left_fold(indexed_results, reduction, initial_value):
    """Reduce from the left in submission order (device_pool.py:246-251)."""
    result = initial_value
    for _, value in sorted(indexed_results, key=lambda item: item[0]):
        result = reduction(result, value)
    return result


def shard_range(total: int, parts: int, index: int) -> tp.Tuple[int, int]:
    """Contiguous shard ``index`` of ``parts``: (first, count) with ceil(total/parts) per shard."""
    per = math.ceil(total / parts) if parts else 0
    first = min(index * per, total)
    return first, min(per, total - first)
