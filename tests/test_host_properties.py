"""Property tests (hypothesis) of the host-side bookkeeping: slices, shards, seeds, batch normalisation."""
import math

from hypothesis import given, settings, strategies as st

from cocos_b200 import distributed as cdist
from cocos_b200.multi_processing import utilities as U
from cocos_b200.numerics import random as rnd


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10_000), batch=st.integers(1, 500))
def test_slices_with_batch_size_partition(n, batch):
    spans = U.generate_slices_with_batch_size(n, batch)
    assert all(0 < e - b <= batch for b, e in spans)
    assert [b for b, _ in spans] == list(range(0, n, batch))
    assert (spans[-1][1] if spans else 0) == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 10_000), parts=st.integers(1, 64))
def test_slices_with_number_of_batches(n, parts):
    spans = U.generate_slices_with_number_of_batches(n, parts)
    assert len(spans) <= parts and spans[0][0] == 0 and spans[-1][1] == n
    assert max(e - b for b, e in spans) == math.ceil(n / parts)


@settings(max_examples=300, deadline=None)
@given(total=st.integers(0, 2 ** 40), parts=st.integers(1, 64))
def test_shard_ranges_tile_the_total(total, parts):
    pos = 0
    for g in range(parts):
        first, count = U.shard_range(total, parts, g)
        assert count >= 0 and (count == 0 or first == pos)
        pos += count
    assert pos == total


@settings(max_examples=200, deadline=None)
@given(R=st.integers(1, 2 ** 40), world=st.sampled_from([1, 2, 3, 4, 8]))
def test_pair_shards_cover_every_pair_once(R, world):
    spans = [cdist.pair_shard(R, world, r) for r in range(world)]
    assert sum(c for _, c in spans) == (R + 1) // 2
    assert all(c <= math.ceil(((R + 1) // 2) / world) for _, c in spans)


@settings(max_examples=100, deadline=None)
@given(seed=st.integers(0, 2 ** 64 - 1), epoch=st.integers(0, 1000), n=st.integers(2, 64))
def test_derived_seeds_are_distinct(seed, epoch, n):
    seeds = {rnd.derive_seed(seed, epoch, i) for i in range(n)}
    assert len(seeds) == n and all(0 <= s < 2 ** 64 for s in seeds)
    assert rnd.derive_seed(seed, epoch, 0) != rnd.derive_seed(seed, epoch + 1, 0)


@settings(max_examples=100, deadline=None)
@given(nb=st.integers(1, 20), with_args=st.booleans(), with_kwargs=st.booleans())
def test_batch_normalisation(nb, with_args, with_kwargs):
    args = [[i] for i in range(nb)] if with_args else None
    kwargs = [{"k": i} for i in range(nb)] if with_kwargs else None
    a, k, n = U._extract_arguments_and_number_of_batches(args, kwargs, None if (with_args or with_kwargs) else nb)
    assert n == nb and len(a) == nb and len(k) == nb
    results = [(i, i) for i in reversed(range(nb))]                        # completion order reversed
    assert U.ordered_left_fold(results, lambda acc, v: acc + [v], []) == list(range(nb))
