"""Heston pricing entry points: examples/stochastic_volatility/heston_utilities.py on B200.

Same names, arguments and return types as the reference:
  simulate_heston_model                         heston_utilities.py:24-96
  compute_option_price_from_simulated_paths     heston_utilities.py:99-119
  simulate_and_compute_option_price             heston_utilities.py:122-164   (the map function)
  simulate_and_compute_option_price_gpu         heston_utilities.py:217-270
  simulate_and_compute_option_price_multicore   heston_utilities.py:167-214   (same estimator, one fused launch)
plus what the fused kernel gives for free: standard errors from antithetic pair
averages, a strike sweep over shared paths (BASELINE config 5), float64 state and the
full log-price trajectory.  The reference's Python time loop (N-1 iterations of ~21
array operations) is replaced by ONE kernel launch.
"""
import ctypes
import math
import typing as tp

import numpy as np

from . import _lib
from .device import ComputeDeviceManager, current_context
from .multi_processing.device_pool import ComputeDevicePool
from .numerics import random as _random
from .numerics._array import empty_on, ndarray
from .numerics.numerical_package_bundle import B200Bundle, NumericalPackageBundle

_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def _require_b200(bundle):
    if bundle is not None and not (isinstance(bundle, type) and issubclass(bundle, B200Bundle)):
        raise TypeError("cocos_b200 runs the Heston path on the B200 bundle only (B200Bundle / CocosBundle); "
                        "the NumPy restatement is test infrastructure under oracle/")


def _params(T, mu, kappa, v_bar, sigma_v, rho, x0, v0) -> "_lib.HestonParams":
    return _lib.HestonParams(float(T), float(mu), float(kappa), float(v_bar), float(sigma_v), float(rho),
                             float(x0), float(v0))


def _suffix(dtype) -> str:
    try:
        return _SUFFIX[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise TypeError(f"dtype {dtype} is not supported (float32 and float64 are)") from None


def simulate_heston_model(T: float, N: int, R: int, mu: float, kappa: float, v_bar: float, sigma_v: float,
                          rho: float, x0: float, v0: float,
                          numerical_package_bundle: tp.Optional[tp.Type[NumericalPackageBundle]] = B200Bundle,
                          dtype=np.float32, return_trajectory: bool = False,
                          return_variance_trajectory: bool = False) -> tp.Tuple:
    """Simulate R antithetic Heston paths over N time points (N-1 Euler-Maruyama steps).

    Returns device arrays (x_T, v_T), each of shape (R,): path i < ceil(R/2) and path
    i + ceil(R/2) are antithetic partners (the reference's layout, random.py:199-214).
    With ``return_trajectory`` a third array of shape (N, 2*ceil(R/2)) holds every
    path's log price at every time point (row 0 = x0; the last column of an odd R is unused); with
    ``return_variance_trajectory`` a fourth array of the same shape holds the variances (row 0 = v0) --
    the path-materialising mode, 8 bytes per path-step written with 2-D TMA tensor stores.
    """
    _require_b200(numerical_package_bundle)
    R, N = int(R), int(N)
    if R < 1:
        raise ValueError("R must be >= 1")
    ctx = current_context()
    sfx = _suffix(dtype)
    pairs = (R + 1) // 2
    # shard layout [pairs | partners]: one contiguous buffer, the first R entries are the reference layout
    x = empty_on(ctx, (2 * pairs,), dtype)
    v = empty_on(ctx, (2 * pairs,), dtype)
    return_trajectory = return_trajectory or return_variance_trajectory
    traj = empty_on(ctx, (N, 2 * pairs), dtype) if return_trajectory else None
    traj_v = empty_on(ctx, (N, 2 * pairs), dtype) if return_variance_trajectory else None
    p = _params(T, mu, kappa, v_bar, sigma_v, rho, x0, v0)
    k_arr, nK = _lib.strikes_array([1.0])
    seed, sub = _random.get_seed(), _random.next_subsequence()
    if traj is None:
        launch = getattr(_lib.lib(), f"cb_heston_price_launch_{sfx}")
        _lib.check(launch(ctx.handle, ctypes.byref(p), R, N, k_arr, nK, seed, sub, 0, pairs, None, x.data_ptr,
                          v.data_ptr, None))
        return x[:R], v[:R]
    launch = getattr(_lib.lib(), f"cb_heston_trajectory_launch_{sfx}")
    _lib.check(launch(ctx.handle, ctypes.byref(p), R, N, k_arr, nK, seed, sub, 0, pairs, None, x.data_ptr, v.data_ptr,
                      traj.data_ptr, traj_v.data_ptr if traj_v is not None else None))
    out = (x[:R], v[:R], traj)
    return out + (traj_v,) if traj_v is not None else out


def call_payoff_sums(x_simulated: ndarray, strikes: tp.Sequence[float]) -> tp.Tuple[np.ndarray, np.ndarray]:
    """(sum of payoffs, sum of squared antithetic-unit averages) per strike, undiscounted."""
    if not isinstance(x_simulated, ndarray):
        raise TypeError("x_simulated must be a cocos_b200 device array (no CPU branch)")
    ctx = _lib.context(x_simulated.device_id)
    k_arr, nK = _lib.strikes_array(strikes)
    sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    fn = getattr(_lib.lib(), f"cb_call_payoff_sums_{_suffix(x_simulated.dtype)}")
    _lib.check(fn(ctx.handle, x_simulated.data_ptr, x_simulated.size, k_arr, nK, sums, sumsq))
    return np.array(sums), np.array(sumsq)


def compute_option_price_from_simulated_paths(
        r: float, T: float, K: float, x_simulated,
        numerical_package_bundle: tp.Optional[tp.Type[NumericalPackageBundle]] = B200Bundle) -> float:
    """exp(-rT) * mean(max(exp(x_T) - K, 0)) over a device array of terminal log prices."""
    _require_b200(numerical_package_bundle)
    sums, _ = call_payoff_sums(x_simulated, [K])
    return float(math.exp(-r * T) * sums[0] / x_simulated.size)


def _price_from_sums(r, T, R, sums, sumsq):
    """Discounted prices and standard errors (from antithetic pair averages) per strike."""
    disc = math.exp(-r * T)
    units = (R + 1) // 2
    sums, sumsq = np.asarray(sums, dtype=np.float64), np.asarray(sumsq, dtype=np.float64)
    price = disc * sums / R
    mean_unit = sums / (2.0 * units) if R % 2 == 0 else sums / R      # odd R: the lone unit is O(1/R)
    var_unit = np.maximum(sumsq / units - mean_unit ** 2, 0.0)
    se = disc * np.sqrt(var_unit / max(units - 1, 1))
    return price, se


SCHEMES = ("euler", "conditional")


def heston_payoff_sums(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, R, strikes, dtype=np.float32,
                       compute_device_pool: tp.Optional[ComputeDevicePool] = None, scheme: str = "euler"):
    """One fused launch per device: undiscounted payoff sums for every strike.

    ``scheme="euler"`` steps log price and variance together, as the reference does.  ``scheme="conditional"``
    (float32, opt-in) samples the log price conditionally on the variance path: the same law of terminal prices and
    of antithetic pairs with half the normals (csrc/heston_conditional.cu), ~1.6x faster.
    """
    if scheme not in SCHEMES:
        raise ValueError(f"scheme must be one of {SCHEMES}")
    R, nT = int(R), int(nT)
    p = _params(T, r, kappa, v_bar, sigma_v, rho, x0, v0)
    seed, sub = _random.get_seed(), _random.next_subsequence()
    sfx = _suffix(dtype)
    if scheme == "conditional" and sfx != "f32":
        raise TypeError("the conditional scheme is float32 only")
    if compute_device_pool is not None:
        return compute_device_pool.heston_sums(p, R, nT, strikes, seed, sub, sfx, scheme=scheme)
    ctx = current_context()
    L = _lib.lib()
    k_arr, nK = _lib.strikes_array(strikes)
    if scheme == "conditional":
        _lib.check(L.cb_heston_price_conditional_launch_f32(ctx.handle, ctypes.byref(p), R, nT, k_arr, nK, seed, sub, 0,
                                                            (R + 1) // 2, None, None, None))
    else:
        launch = getattr(L, f"cb_heston_price_launch_{sfx}")
        _lib.check(launch(ctx.handle, ctypes.byref(p), R, nT, k_arr, nK, seed, sub, 0, (R + 1) // 2, None, None, None,
                          None))
    sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    _lib.check(L.cb_heston_price_finish(ctx.handle, nK, sums, sumsq))
    return list(sums), list(sumsq)


def simulate_and_compute_option_price(
        x0: float, v0: float, r: float, rho: float, sigma_v: float, kappa: float, v_bar: float, T: float, K: float,
        nT: int, R: int, numerical_package_bundle: tp.Optional[tp.Type[NumericalPackageBundle]] = B200Bundle,
        verbose: bool = False, scheme: str = "euler") -> float:
    """Price a European call under Heston by Monte Carlo: the fused kernel, one launch."""
    _require_b200(numerical_package_bundle)
    if verbose:
        print(f'computing on device={ComputeDeviceManager.get_current_compute_device_id()}')
    sums, _ = heston_payoff_sums(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, R, [K], scheme=scheme)
    return float(math.exp(-r * T) * sums[0] / int(R))


def price_with_standard_error(x0, v0, r, rho, sigma_v, kappa, v_bar, T, K, nT, R, dtype=np.float32,
                              compute_device_pool: tp.Optional[ComputeDevicePool] = None,
                              scheme: str = "euler") -> tp.Tuple[float, float]:
    sums, sumsq = heston_payoff_sums(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, R, [K], dtype, compute_device_pool,
                                     scheme)
    price, se = _price_from_sums(r, T, int(R), sums, sumsq)
    return float(price[0]), float(se[0])


def price_strike_sweep(x0, v0, r, rho, sigma_v, kappa, v_bar, T, strikes, nT, R, dtype=np.float64,
                       compute_device_pool: tp.Optional[ComputeDevicePool] = None, scheme: str = "euler"):
    """Prices and standard errors for up to 64 strikes sharing one set of paths (BASELINE config 5)."""
    sums, sumsq = heston_payoff_sums(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, R, strikes, dtype,
                                     compute_device_pool, scheme)
    return _price_from_sums(r, T, int(R), sums, sumsq)


def simulate_and_compute_option_price_gpu(
        x0: float, v0: float, r: float, rho: float, sigma_v: float, kappa: float, v_bar: float, T: float, K: float,
        nT: int, R: int, compute_device_pool: tp.Optional[ComputeDevicePool] = None,
        number_of_batches: tp.Optional[int] = None, verbose: bool = False) -> float:
    """Multi-GPU pricing (heston_utilities.py:217-270).

    The reference runs ``number_of_batches`` independent simulations of ceil(R/nb)
    paths through ``map_reduce`` and averages the batch prices, i.e. it prices with
    nb*ceil(R/nb) paths in total.  Here that total is simulated as ONE sharded run:
    each GPU takes a contiguous block of antithetic pairs and a single NCCL allreduce
    combines the sums -- the same estimator up to floating-point summation order.
    """
    if number_of_batches is None:
        number_of_batches = 1 if compute_device_pool is None else compute_device_pool.number_of_devices
    if compute_device_pool is None:
        if verbose:
            print(f'computing {R} paths on single GPU')
        return simulate_and_compute_option_price(x0=x0, v0=v0, r=r, rho=rho, sigma_v=sigma_v, kappa=kappa,
                                                 v_bar=v_bar, T=T, K=K, nT=nT, R=R)
    per_batch = math.ceil(R / number_of_batches)
    total = per_batch * number_of_batches
    if verbose:
        print(f'computing {R} paths on {compute_device_pool.number_of_devices} GPUs in '
              f'{number_of_batches} batches of {per_batch} paths')
    sums, _ = heston_payoff_sums(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, total, [K],
                                 compute_device_pool=compute_device_pool)
    return float(math.exp(-r * T) * sums[0] / total)


def simulate_and_compute_option_price_multicore(
        x0: float, v0: float, r: float, rho: float, sigma_v: float, kappa: float, v_bar: float, T: float, K: float,
        nT: int, R: int, numerical_package_bundle: tp.Optional[tp.Type[NumericalPackageBundle]] = B200Bundle,
        number_of_cores: tp.Optional[int] = None) -> float:
    """The reference's CPU multi-core entry (heston_utilities.py:167-214) maps ``simulate_and_compute_option_price``
    on its NumPy bundle over ``number_of_cores`` (default ``5 * cpu_count()``) batches of ``ceil(R / number_of_cores)``
    paths and averages the batch prices.  There is no CPU branch here (a NumPy bundle is refused like everywhere
    else); with the B200 bundle the same estimator -- ``number_of_cores * ceil(R / number_of_cores)`` paths in all --
    runs as ONE fused launch, which is what the batches were a workaround for."""
    _require_b200(numerical_package_bundle)
    if number_of_cores is None:
        import multiprocessing
        number_of_cores = 5 * multiprocessing.cpu_count()
    number_of_cores = int(number_of_cores)
    if number_of_cores < 1:
        raise ValueError("number_of_cores must be at least 1")
    total = math.ceil(int(R) / number_of_cores) * number_of_cores
    return simulate_and_compute_option_price(x0=x0, v0=v0, r=r, rho=rho, sigma_v=sigma_v, kappa=kappa, v_bar=v_bar,
                                             T=T, K=K, nT=nT, R=total)


PAYOFF_KINDS = {"european": 0, "asian": 1, "up_and_out": 2, "down_and_out": 3}


def price_path_dependent(x0, v0, r, rho, sigma_v, kappa, v_bar, T, K, nT, R, kind: str = "asian", option: str = "call",
                         barrier: tp.Optional[float] = None, dtype=np.float32,
                         compute_device_pool: tp.Optional[ComputeDevicePool] = None):
    """Price and standard error of a path-dependent option on the simulated Heston paths (one fused launch).

    kind: "european", "asian" (arithmetic average of the nT-1 simulated prices), "up_and_out" / "down_and_out"
    (knock-out monitored at the simulated dates, ``barrier`` is a price level); option: "call" or "put";
    ``K`` may be a sequence of up to 64 strikes sharing the paths (arrays are returned then).
    The reference only motivates these payoffs (README.md:376-382); the path functionals live in registers.
    """
    if kind not in PAYOFF_KINDS:
        raise ValueError(f"kind must be one of {tuple(PAYOFF_KINDS)}")
    if option not in ("call", "put"):
        raise ValueError("option must be 'call' or 'put'")
    if PAYOFF_KINDS[kind] >= 2 and (barrier is None or barrier <= 0):
        raise ValueError("a positive barrier level is required for knock-out payoffs")
    strikes = list(K) if isinstance(K, (list, tuple, np.ndarray)) else [K]
    R, nT = int(R), int(nT)
    p = _params(T, r, kappa, v_bar, sigma_v, rho, x0, v0)
    spec = _lib.PayoffSpec(PAYOFF_KINDS[kind], int(option == "put"), float(barrier or 0.0))
    seed, sub = _random.get_seed(), _random.next_subsequence()
    sfx = _suffix(dtype)
    if compute_device_pool is None:
        ctx = current_context()
        L = _lib.lib()
        k_arr, nK = _lib.strikes_array(strikes)
        launch = getattr(L, f"cb_heston_payoff_launch_{sfx}")
        _lib.check(launch(ctx.handle, ctypes.byref(p), R, nT, k_arr, nK, seed, sub, 0, (R + 1) // 2, None,
                          ctypes.byref(spec)))
        sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
        _lib.check(L.cb_heston_price_finish(ctx.handle, nK, sums, sumsq))
    else:
        sums, sumsq = compute_device_pool.heston_sums(p, R, nT, strikes, seed, sub, sfx, spec=spec)
    price, se = _price_from_sums(r, T, R, list(sums), list(sumsq))
    if isinstance(K, (list, tuple, np.ndarray)):
        return price, se
    return float(price[0]), float(se[0])
