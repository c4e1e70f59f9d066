"""Monte-Carlo pi (BASELINE config 2), the payoff reduction on arrays, and ComputeDevicePool."""
import ctypes
import math

import numpy as np
import pytest

from gpu_util import fused, pi_count

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gpu_device):
    return gpu_device.current_context()


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 1_000_001, 30_000_000])
@pytest.mark.parametrize("antithetic", [False, True])
def test_pi_count_bit_exact(ctx, n, antithetic):
    from oracle import c_oracle
    assert pi_count(ctx, n, seed=17, sub=2, antithetic=antithetic) == c_oracle.pi_inside(n, 17, 2, antithetic)


def test_pi_draw_shards_add_up(ctx):
    n = 10_000_001
    ndraw = (n + 1) // 2
    whole = pi_count(ctx, n, 3, 0, True)
    cuts = [0, 1, 7, ndraw // 3, ndraw // 3 + 1, ndraw]
    parts = sum(pi_count(ctx, n, 3, 0, True, first=a, count=b - a) for a, b in zip(cuts[:-1], cuts[1:]))
    assert parts == whole


def test_estimate_pi_config2(gpu_device):
    """BASELINE config 2: 1e9 antithetic uniform draws on one GPU."""
    from cocos_b200 import monte_carlo_pi as mcp
    from cocos_b200.numerics import random as rnd
    rnd.seed(0)
    est = mcp.estimate_pi(1_000_000_000, batches=1, antithetic=True)
    assert abs(est - math.pi) < 5 * 1.64 / math.sqrt(1e9) + 1e-4      # sd of 4*Bernoulli(pi/4) is 1.64
    rnd.seed(0)
    est20 = mcp.estimate_pi(1_000_000_000, batches=20)               # the shipped example: 20 plain batches
    assert abs(est20 - math.pi) < 5 * 1.64 / math.sqrt(1e9) + 1e-4
    with pytest.raises(NotImplementedError):
        mcp.estimate_pi(10, gpu=False)
    from cocos_b200.multi_processing import ComputeDevicePool
    from oracle import c_oracle
    pool = ComputeDevicePool(compute_devices=[0])                    # pool path on one device
    rnd.seed(3)
    assert mcp.count_inside(1_000_001, True, compute_device_pool=pool) == c_oracle.pi_inside(1_000_001, 3, 0, True)
    pool.close()


def test_payoff_from_array_matches_reference_price(gpu_device, golden_heston, ref_params):
    """compute_option_price_from_simulated_paths on the reference's own terminal values."""
    from cocos_b200 import heston
    from gpu_util import to_device
    ctx = gpu_device.current_context()
    for name in ("even", "odd", "one", "large"):
        x = golden_heston[name + "_x"]
        want = float(golden_heston[name + "_price"])
        got = heston.compute_option_price_from_simulated_paths(ref_params["r"], ref_params["T"], ref_params["K"],
                                                               to_device(ctx, x))
        assert got == pytest.approx(want, rel=2e-6)        # float64 accumulation vs the reference's float32 mean
    from oracle import c_oracle
    x = golden_heston["odd_x"].astype(np.float64)
    s, q = heston.call_payoff_sums(to_device(ctx, x), [0.8, 0.95, 1.1])
    ws, wq = c_oracle.payoff_sums(x, [0.8, 0.95, 1.1])
    np.testing.assert_allclose(s, ws, rtol=1e-13)
    np.testing.assert_allclose(q, wq, rtol=1e-13)


def test_simulate_heston_model_surface(gpu_device, ref_params):
    from cocos_b200 import heston, B200Bundle, CocosBundle
    from cocos_b200.numerics import random as rnd
    q = ref_params
    rnd.seed(4)
    x, v = heston.simulate_heston_model(T=q["T"], N=51, R=10_001, mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"],
                                        sigma_v=q["sigma_v"], rho=q["rho"], x0=q["x0"], v0=q["v0"],
                                        numerical_package_bundle=CocosBundle)
    assert x.shape == (10_001,) and v.shape == (10_001,) and x.dtype == np.float32
    price = heston.compute_option_price_from_simulated_paths(q["r"], q["T"], q["K"], x, B200Bundle)
    rnd.seed(4)
    fusedp = heston.simulate_and_compute_option_price(nT=51, R=10_001, numerical_package_bundle=B200Bundle, **q)
    assert price == pytest.approx(fusedp, rel=2e-6)
    B200Bundle.synchronize()
    assert B200Bundle.module().random is B200Bundle.random_module()
    with pytest.raises(NotImplementedError):
        _ = x + x


def test_pool_generic_map_reduce(gpu_device):
    """Ordered left fold, argument normalisation, transfer hooks: device_pool.py:150-253 semantics."""
    from cocos_b200.multi_processing import ComputeDevicePool
    pool = ComputeDevicePool()
    assert pool.number_of_devices >= 1
    order = pool.map_reduce(f=lambda tag: tag, reduction=lambda acc, t: acc + [t], initial_value=[],
                            args_list=[[i] for i in range(7)])
    assert order == list(range(7))
    total = pool.map_reduce(f=lambda a, b=0: a * b, reduction=lambda x, y: x + y, initial_value=0.0,
                            host_to_device_transfer_function=lambda a, b=0: ((a + 1,), dict(b=b)),
                            device_to_host_transfer_function=lambda r: 10 * r,
                            args_list=[[1], [2]], kwargs_list=[dict(b=3), dict(b=4)])
    assert total == 10 * (2 * 3 + 3 * 4)
    assert pool.map_reduce(f=lambda: 2.0, reduction=lambda x, y: x + y / 4, initial_value=0.0,
                           number_of_batches=4) == 2.0
    with pytest.raises(ValueError):
        pool.map_reduce(f=lambda: 1, reduction=lambda x, y: x + y, initial_value=0)
    combined = pool.map_combine(f=lambda i: [i, i], combination=lambda rs: sum(rs, []), args_list=[[1], [2], [3]])
    assert combined == [1, 1, 2, 2, 3, 3]
    with pytest.raises(TypeError):
        ComputeDevicePool(compute_devices=["gpu0"])
    pool.sync()


def test_pool_heston_reference_call_pattern(gpu_device, golden_values, ref_params):
    """The reference's own multi-GPU call: map_reduce(f=simulate_and_compute_option_price, lambda fold)."""
    from cocos_b200 import heston
    from cocos_b200.multi_processing import ComputeDevicePool
    from cocos_b200.numerics import random as rnd
    pool = ComputeDevicePool()
    nb, R = 4, 1_000_000
    kwargs = dict(ref_params, nT=101, R=math.ceil(R / nb))
    rnd.seed(9)
    via_map = pool.map_reduce(f=heston.simulate_and_compute_option_price,
                              reduction=lambda x, y: x + y / nb, initial_value=0.0, kwargs_list=nb * [kwargs])
    rnd.seed(9)
    again = pool.map_reduce(f=heston.simulate_and_compute_option_price,
                            reduction=lambda x, y: x + y / nb, initial_value=0.0, kwargs_list=nb * [kwargs])
    assert via_map == again                                  # per-batch derived seeds: scheduling-independent
    ref = golden_values["statistical"]["nT101"]
    assert abs(via_map - ref["price"]) <= 3 * math.hypot(ref["se"], ref["se"]) + 1e-4
    rnd.seed(9)
    sharded = heston.simulate_and_compute_option_price_gpu(nT=101, R=R, compute_device_pool=pool,
                                                           number_of_batches=nb, **ref_params)
    assert abs(sharded - ref["price"]) <= 3 * math.hypot(ref["se"], ref["se"]) + 1e-4
    single = heston.simulate_and_compute_option_price_gpu(nT=101, R=R, **ref_params)
    assert abs(single - ref["price"]) <= 3 * math.hypot(ref["se"], ref["se"]) + 1e-4
    pool.close()


def test_allreduce_path_on_one_rank(gpu_device, ref_params):
    """The NCCL allreduce that follows the kernel, exercised with a 1-rank communicator."""
    from cocos_b200 import _lib
    ctx = gpu_device.current_context()
    q = ref_params
    P = dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
             x0=q["x0"], v0=q["v0"])
    ctxs = (ctypes.c_void_p * 1)(ctx.handle)
    comms = (ctypes.c_void_p * 1)()
    _lib.check(_lib.lib().cb_comm_init_all(1, ctxs, comms))
    comm = ctypes.c_void_p(comms[0])
    try:
        a = fused(ctx, P, 200_000, 33, [0.9, 1.0], seed=2, sub=1)
        b = fused(ctx, P, 200_000, 33, [0.9, 1.0], seed=2, sub=1, comm=comm)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    finally:
        _lib.lib().cb_comm_destroy(comm)


def test_array_roundtrip(gpu_device):
    import cocos_b200.numerics as cn
    a = np.arange(12, dtype=np.float64).reshape(3, 4)
    d = cn.array(a)
    assert d.shape == (3, 4) and d.dtype == np.float64 and np.array_equal(np.array(d), a)
    assert np.array_equal(cn.asnumpy(d[1]), a[1])
    f = d.astype(np.float32)
    assert f.dtype == np.float32 and np.array_equal(f.to_numpy(), a.astype(np.float32))
    assert cn.array([1.0, 2.0]).dtype == np.float32 and len(cn.array([1.0, 2.0])) == 2
    assert float(cn.array(np.float32(2.5))) == 2.5


def test_pi_window_across_the_32_bit_block_boundary(ctx):
    """Draw windows whose Philox block indices straddle 2^32 (the kernel cuts the index range into segments with a
    uniform high counter word): the count equals the one computed on the host from the raw Philox words of the same
    blocks (cb_philox_words), for plain and antithetic points, and windows add up across the boundary."""
    from cocos_b200 import _lib
    from cocos_b200.numerics._array import empty_on
    n = 1 << 36                                      # 6.9e10 points: only windows of it are evaluated
    seed, sub = 29, 6
    boundary_draw = 2 * (1 << 32)                    # first draw of block 2^32
    lo, hi = boundary_draw - 1001, boundary_draw + 1503
    for antithetic in (False, True):
        got = pi_count(ctx, n, seed, sub, antithetic, first=lo, count=hi - lo)
        a = pi_count(ctx, n, seed, sub, antithetic, first=lo, count=boundary_draw - 3 - lo)
        b = pi_count(ctx, n, seed, sub, antithetic, first=boundary_draw - 3, count=hi - (boundary_draw - 3))
        assert a + b == got
        # host evaluation from the raw words of blocks [lo//2, (hi+1)//2)
        b0, b1 = lo // 2, (hi + 1) // 2
        words = empty_on(ctx, (b1 - b0, 4), np.uint32)
        _lib.check(_lib.lib().cb_philox_words(ctx.handle, words.data_ptr, b1 - b0, b0, 0, seed, sub))
        w = words.to_numpy()
        u = ((w >> np.uint32(9)).astype(np.uint32) + np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
        ux = u[:, [0, 2]].reshape(-1)                 # draw 2b+h uses words (2h, 2h+1) of block b
        uy = u[:, [1, 3]].reshape(-1)
        j = np.arange(2 * b0, 2 * b1)
        own = (j >= lo) & (j < hi)
        inside = (ux * ux + uy * uy) <= np.float32(1.0)
        expect = int(np.count_nonzero(inside & own))
        if antithetic:
            rx, ry = np.float32(1.0) - ux, np.float32(1.0) - uy
            expect += int(np.count_nonzero(((rx * rx + ry * ry) <= np.float32(1.0)) & own & (j < n // 2)))
        assert got == expect
