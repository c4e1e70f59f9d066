"""Host-side mirror of the reference interface: pure-Python pieces, no GPU needed."""
import math
import os

import numpy as np
import pytest

from conftest import ROOT
from cocos_b200.multi_processing import utilities as U
from cocos_b200.numerics import random as rnd


def test_antithetic_slices_reference_cases():
    """cocos/tests/test_numerics/test_random/test_antithetic_slices.py:4-15, verbatim expectations."""
    assert rnd.get_antithetic_slices((2, 10), antithetic_dimension=0) == (slice(0, 1, 1), slice(0, 10, 1))
    assert rnd.get_antithetic_slices((2, 10), antithetic_dimension=1) == (slice(0, 2, 1), slice(0, 5, 1))
    assert rnd.get_antithetic_slices((2, 9), antithetic_dimension=1) == (slice(0, 2, 1), slice(0, 4, 1))


@pytest.mark.parametrize('n, batch_size, result', [(10, 2, ((0, 2), (2, 4), (4, 6), (6, 8), (8, 10))),
                                                   (10, 3, ((0, 3), (3, 6), (6, 9), (9, 10))),
                                                   (10, 4, ((0, 4), (4, 8), (8, 10)))])
def test_generate_slices_with_batch_size(n, batch_size, result):
    """cocos/tests/test_multiprocessing/test_generate_slices.py:10-13."""
    assert result == U.generate_slices_with_batch_size(n=n, batch_size=batch_size)


@pytest.mark.parametrize('n, number_of_batches, result', [(10, 2, ((0, 5), (5, 10))),
                                                          (10, 3, ((0, 4), (4, 8), (8, 10))),
                                                          (10, 4, ((0, 3), (3, 6), (6, 9), (9, 10)))])
def test_generate_slices_with_number_of_batches(n, number_of_batches, result):
    """cocos/tests/test_multiprocessing/test_generate_slices.py:16-20."""
    assert result == U.generate_slices_with_number_of_batches(n=n, number_of_batches=number_of_batches)


def test_helpers_against_golden(golden_values):
    k = golden_values["helper_kats"]
    for case in k["antithetic_slices"]:
        got = rnd.get_antithetic_slices(case["shape"], case["axis"])
        assert [[s.start, s.stop, s.step] for s in got] == case["expect"]
    a, kw, n = U._extract_arguments_and_number_of_batches(kwargs_list=[{"q": 1}, {"q": 2}])
    assert ([list(x) for x in a], kw, n) == (k["extract_kwargs_only"]["args"], k["extract_kwargs_only"]["kwargs"], 2)
    a, kw, n = U._extract_arguments_and_number_of_batches(args_list=[[1], [2], [3]])
    assert ([list(x) for x in a], kw, n) == (k["extract_args_only"]["args"], k["extract_args_only"]["kwargs"], 3)
    a, kw, n = U._extract_arguments_and_number_of_batches(number_of_batches=4)
    assert ([list(x) for x in a], kw, n) == (k["extract_count_only"]["args"], k["extract_count_only"]["kwargs"], 4)
    with pytest.raises(ValueError):
        U._extract_arguments_and_number_of_batches()


def test_shape_and_axis_errors():
    with pytest.raises(ValueError):
        rnd.verify_shape_and_antithetic_dimension((1, 2, 3, 4, 5), 0)
    with pytest.raises(ValueError):
        rnd.verify_shape_and_antithetic_dimension((), 0)
    with pytest.raises(ValueError):
        rnd.verify_shape_and_antithetic_dimension((3, 3), 2)
    with pytest.raises(ValueError):
        rnd.verify_shape_and_antithetic_dimension((3, 3), -1)
    with pytest.raises(TypeError):                       # the reference's declared default (None) fails the same way
        rnd.verify_shape_and_antithetic_dimension((3, 3), None)
    rnd.verify_shape_and_antithetic_dimension((3, 3), 1)


def test_ordered_left_fold():
    results = [(2, "c"), (0, "a"), (1, "b")]             # completion order != submission order
    assert U.ordered_left_fold(results, lambda acc, x: acc + x, "") == "abc"
    nb = 4
    assert U.ordered_left_fold([(i, 1.0) for i in range(nb)], lambda x, y: x + y / nb, 0.0) == 1.0


@pytest.mark.parametrize("total,parts", [(0, 3), (1, 8), (7, 2), (10, 3), (5_000_000, 8), (500_000_001, 8)])
def test_shard_range_partitions(total, parts):
    covered = 0
    for g in range(parts):
        first, count = U.shard_range(total, parts, g)
        assert first == covered or count == 0
        assert count <= math.ceil(total / parts) if total else count == 0
        covered += count
    assert covered == total


def test_generator_state():
    rnd.seed(None)
    assert rnd.get_seed() == 0 and rnd.get_state() == (0, 0)
    rnd.seed(2 ** 64 + 5)
    assert rnd.get_seed() == 5
    assert [rnd.next_subsequence() for _ in range(3)] == [0, 1, 2]
    rnd.seed(5)
    assert rnd.get_state() == (5, 0)
    seeds = {rnd.derive_seed(5, e, i) for e in range(4) for i in range(64)}
    assert len(seeds) == 256 and all(0 <= s < 2 ** 64 for s in seeds)
    assert rnd.derive_seed(5, 1, 2) == rnd.derive_seed(5, 1, 2) != rnd.derive_seed(6, 1, 2)


def test_price_from_sums_matches_oracle():
    from cocos_b200.heston import _price_from_sums
    from oracle import heston_numpy as hn
    rng = np.random.default_rng(0)
    for R in (10, 11, 5001):
        x = rng.normal(0.0, 0.2, R)
        s, q = hn.payoff_sums(x, [0.9, 1.0])
        price, se = _price_from_sums(0.03, 1.0, R, s, q)
        for j, K in enumerate((0.9, 1.0)):
            want_p, want_se = hn.call_price_and_se(0.03, 1.0, K, x)
            assert price[j] == pytest.approx(want_p, rel=1e-12)
            assert se[j] == pytest.approx(want_se, rel=3.0 / R if R % 2 else 1e-9)   # odd R: the lone path is O(1/R)


def test_map_reduce_multicore_semantics():
    from cocos_b200.multi_processing import map_combine_multicore, map_reduce_multicore
    nb = 5
    total = map_reduce_multicore(f=lambda a, b=1: a * b, reduction=lambda x, y: x + y / nb, initial_value=0.0,
                                 args_list=[[i] for i in range(nb)], kwargs_list=nb * [dict(b=2)])
    assert total == pytest.approx(sum(2 * i for i in range(nb)) / nb)
    order = map_combine_multicore(f=lambda i: i * i, combination=list, args_list=[[i] for i in range(6)])
    assert order == [i * i for i in range(6)]
    assert map_reduce_multicore(f=lambda: 3, reduction=lambda x, y: x + y, initial_value=0, number_of_batches=3) == 9
    with pytest.raises(ValueError):
        map_reduce_multicore(f=lambda: 3, reduction=lambda x, y: x + y, initial_value=0)


def test_bundle_protocol():
    from cocos_b200 import B200Bundle, CocosBundle, NumericalPackageBundle
    import cocos_b200.numerics as cn
    assert CocosBundle is B200Bundle and issubclass(B200Bundle, NumericalPackageBundle)
    assert B200Bundle.is_installed() and B200Bundle.label() == 'Cocos-B200'
    assert B200Bundle.module() is cn and B200Bundle.random_module() is cn.random
    for name in ("rand", "randn", "rand_with_dtype", "randn_with_dtype", "randn_antithetic", "rand_antithetic",
                 "get_antithetic_slices", "seed", "get_seed"):
        assert callable(getattr(cn.random, name))
    from cocos_b200.heston import _require_b200
    with pytest.raises(TypeError):
        _require_b200(int)
    from cocos_b200.numerics.numerical_package_selector import select_num_pack
    assert select_num_pack(True) is cn
    with pytest.raises(NotImplementedError):
        select_num_pack(False)


def test_bundled_nccl_is_found_without_importing_torch():
    """The ctypes layer names torch's own libnccl to the C library before any communicator exists (a later
    `import torch` must not find an older system copy resident under the same SONAME)."""
    import subprocess
    import sys
    code = ("import sys, os\n"
            "from cocos_b200 import _lib\n"
            "p = _lib._bundled_nccl()\n"
            "assert 'torch' not in sys.modules\n"
            "assert p is None or (os.path.isabs(p) and os.path.exists(p) and p.endswith('libnccl.so.2')), p\n"
            "print('ok', p)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stderr[-1500:]


def test_reference_staging_is_byte_exact_and_tamper_evident(tmp_path, monkeypatch):
    """oracle/build_ref.py: the staged copy the bench's CPU arm imports is the reference byte for byte (SHA-256
    manifest), and an edited staged file fails verification."""
    from oracle import build_ref
    if not build_ref.source_available():
        pytest.skip("/root/reference is not mounted here")
    monkeypatch.setattr(build_ref, "STAGED_ROOT", str(tmp_path / "_ref"))
    monkeypatch.setattr(build_ref, "MANIFEST", str(tmp_path / "_ref" / "MANIFEST.json"))
    assert build_ref.stage(verbose=False) and build_ref.verify()
    staged = tmp_path / "_ref" / "examples" / "stochastic_volatility" / "heston_utilities.py"
    source = os.path.join(build_ref.SOURCE_ROOT, "examples", "stochastic_volatility", "heston_utilities.py")
    assert staged.read_bytes() == open(source, "rb").read()
    assert not (tmp_path / "_ref" / "cocos" / "tests").exists()          # the reference's tests are not staged
    staged.write_text(staged.read_text() + "\n# edited\n")
    assert not build_ref.verify()
    assert build_ref.stage(verbose=False) and build_ref.verify()         # staging again restores the original


def test_bench_reads_roofline_traffic_from_the_committed_profile():
    """bench.py's roofline.traffic is dram read + write bytes per launch from the ncu summary under profiles/."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    traffic, source = bench.profiled_traffic()
    assert source is not None and source.startswith("profiles/") and os.path.exists(os.path.join(ROOT, source))
    assert traffic is not None and 0 <= traffic < 10_000_000             # a compute-bound kernel: kilobytes, not gigabytes


def test_integer_widening_of_negative_floats_is_exact():
    """csrc/heston_step.cuh widen_negative<double>: a NORMAL, strictly negative float is widened to double by
    re-biasing the exponent with integer arithmetic (hi = (bits >> 3) + 0xA8000000, lo = bits << 29) instead of F2F.
    The same arithmetic in NumPy must reproduce astype(float64) bit for bit, from -FLT_MIN to -FLT_MAX."""
    rng = np.random.default_rng(5)
    mant = rng.integers(0, 1 << 23, size=1_000_000, dtype=np.uint32)
    expo = rng.integers(1, 255, size=mant.size, dtype=np.uint32)              # normal numbers only
    bits = (np.uint32(1) << np.uint32(31)) | (expo << np.uint32(23)) | mant
    bits = np.concatenate([bits, np.array([0x80800000, 0xFF7FFFFF, 0xBF800000, 0xB3800000], dtype=np.uint32)])
    hi = ((bits >> np.uint32(3)).astype(np.uint64) + np.uint64(0xA8000000)) & np.uint64(0xFFFFFFFF)
    lo = (bits.astype(np.uint64) << np.uint64(29)) & np.uint64(0xFFFFFFFF)
    widened = ((hi << np.uint64(32)) | lo).view(np.float64)
    assert np.array_equal(widened, bits.view(np.float32).astype(np.float64))
    assert np.all(widened < 0)
