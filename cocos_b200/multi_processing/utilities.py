"""Batch bookkeeping shared by the map-reduce front ends (reference: cocos/multi_processing/utilities.py)."""
import enum
import typing as tp

ResultType = tp.TypeVar('ResultType')
ParameterTransferFunction = tp.Callable[..., tp.Tuple[tp.Sequence, tp.Dict[str, tp.Any]]]
Span = tp.Tuple[int, int]


class MultiprocessingPoolType(enum.Enum):
    """The reference chooses between loky and pathos process pools (utilities.py:23-40).  Here the device pool
    lives in one process (streams + NCCL) and the CPU pool is a concurrent.futures executor; the two reference
    members are kept as names only."""
    LOKY = 1
    PATHOS = 2
    STREAMS = 3
    FUTURES = 4

    @staticmethod
    def default() -> 'MultiprocessingPoolType':
        return MultiprocessingPoolType.STREAMS


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def generate_slices_with_batch_size(n: int, batch_size: int) -> tp.Tuple[Span, ...]:
    """Cut range(n) into consecutive (begin, end) spans no longer than ``batch_size`` (utilities.py:43-63)."""
    if batch_size <= 0:
        raise ValueError("batch_size must be positive")
    return tuple((begin, min(begin + batch_size, n)) for begin in range(0, n, batch_size))


def generate_slices_with_number_of_batches(n: int, number_of_batches: int) -> tp.Tuple[Span, ...]:
    """Spans of length ceil(n / number_of_batches); the last one may be shorter (utilities.py:66-79)."""
    return generate_slices_with_batch_size(n, _ceil_div(n, number_of_batches))


def _extract_arguments_and_number_of_batches(
        args_list: tp.Optional[tp.Sequence[tp.Sequence]] = None,
        kwargs_list: tp.Optional[tp.Sequence[tp.Dict[str, tp.Any]]] = None,
        number_of_batches: tp.Optional[int] = None):
    """Normalise the three optional batch descriptions of ``map_reduce`` / ``map_combine`` (utilities.py:88-106).

    The batch count comes from ``number_of_batches``, else from ``args_list``, else from ``kwargs_list``; a missing
    list becomes that many empty argument lists / dicts.  With nothing to go by the call is an error.
    """
    if number_of_batches is None:
        given = next((lst for lst in (args_list, kwargs_list) if lst is not None), None)
        if given is None:
            raise ValueError('Number_of_batches must be defined if both args_list and kwargs_list are empty')
        number_of_batches = len(given)
    positional = args_list if args_list is not None else [[] for _ in range(number_of_batches)]
    keyword = kwargs_list if kwargs_list is not None else [{} for _ in range(number_of_batches)]
    return positional, keyword, number_of_batches


def ordered_left_fold(indexed_results, reduction, initial_value):
    """Fold (index, value) pairs from the left in INDEX order, whatever order they completed in
    (device_pool.py:246-251)."""
    accumulator = initial_value
    for _, value in sorted(indexed_results, key=lambda item: item[0]):
        accumulator = reduction(accumulator, value)
    return accumulator


def shard_range(total: int, parts: int, index: int) -> Span:
    """Contiguous shard ``index`` of ``parts``: (first, count) with ceil(total/parts) items per shard."""
    per_shard = _ceil_div(total, parts) if parts else 0
    first = min(index * per_shard, total)
    return first, min(per_shard, total - first)
