"""Gaussian kernel density estimation on B200: the evaluation part of cocos/scientific/kde.py.

Mirrors ``gaussian_kde`` (kde.py:414-1170, itself adapted from SciPy): constructor arguments, ``d / n / factor /
covariance / inv_cov / weights / neff``, ``evaluate`` (= ``__call__`` = ``pdf``), ``logpdf``,
``evaluate_in_batches``, ``evaluate_in_batches_on_multiple_devices``, ``integrate_gaussian``, ``integrate_box_1d``,
``integrate_kde``, ``resample`` (kde.py:816-1025) and the module-level ``evaluate_gaussian_kde_in_batches``
(kde.py:1205-1232).

The reference whitens the data with cholesky(inv_cov) and evaluates
``norm * sum_i w_i exp(-|p_i - q_j|^2 / 2)`` by materialising an (n, m, d) array per batch (kde.py:152-197).  Here the
d x d algebra (covariance, inverse, Cholesky: parameter preparation) stays on the host, the data moments, the whitening
and the O(n*m) sum are CUDA kernels (csrc/kde_kernels.cu); batching only survives as API.
"""
import ctypes
import math
import numbers
import typing as tp

import numpy as np

from .. import _lib
from ..device import ComputeDeviceManager, current_context
from ..multi_processing.device_pool import ComputeDevicePool
from ..multi_processing.single_device_batch_processing import map_combine_single_device
from ..multi_processing.utilities import generate_slices_with_number_of_batches
from ..numerics._array import empty_on, ndarray

SCOTTS_FACTOR_STRING = 'scotts'
SILVERMAN_FACTOR_STRING = 'silverman'
_PRESCALE = math.sqrt(0.5 * math.log2(math.e))       # exp(-r^2/2) = 2^(-(s r)^2)


class GaussianKDEInformation:
    """What a bandwidth rule may look at (kde.py:363-369)."""

    def __init__(self, dimension: int, n: int, neff: float, points, weights):
        self.dimension, self.n, self.neff, self.points, self.weights = dimension, n, neff, points, weights


def compute_scotts_factor(kde_info: GaussianKDEInformation) -> float:
    return np.power(kde_info.neff, -1.0 / (kde_info.dimension + 4))


def compute_silverman_factor(kde_info: GaussianKDEInformation) -> float:
    d, neff = kde_info.dimension, kde_info.neff
    return np.power(neff * (d + 2.0) / 4.0, -1.0 / (d + 4))


def _to_device_rows(ctx, array, d_expected: tp.Optional[int] = None) -> ndarray:
    """(d, n) or (n,) input (NumPy or device) -> contiguous float32 device array of shape (n, d)."""
    if isinstance(array, ndarray):
        t = array.reshape(1, -1) if array.ndim == 1 else array
        if t.ndim != 2:
            raise ValueError("the dataset must be a (d, n) array or an (n,) vector")
        if t.device_id != ctx.device:
            t = t.copy_to(ctx)
        if t.dtype != np.dtype(np.float32):
            t = t.astype(np.float32)
        out = t.reshape(t.shape[1], 1) if t.shape[0] == 1 else t.transpose()        # (d, n) -> (n, d)
    else:
        host = np.atleast_2d(np.asarray(array, dtype=np.float32))
        rows_h = np.ascontiguousarray(host.T)
        out = empty_on(ctx, rows_h.shape, np.float32)
        _lib.check(_lib.lib().cb_memcpy_h2d(ctx.handle, out.data_ptr, rows_h.ctypes.data, rows_h.nbytes))
    if d_expected is not None and out.shape[1] != d_expected:
        raise ValueError(f"points have dimension {out.shape[1]}, dataset has dimension {d_expected}")
    return out


class gaussian_kde:
    def __init__(self, dataset, bw_method: tp.Optional[tp.Union[str, tp.Callable, numbers.Number]] = None,
                 weights=None, gpu: bool = True):
        if not gpu:
            raise NotImplementedError("cocos_b200 has no CPU branch; use scipy.stats.gaussian_kde on the host")
        self._gpu = True
        ctx = current_context()
        self._ctx = ctx
        self._points = _to_device_rows(ctx, dataset)                   # (n, d) float32 on the device
        self.n, self.d = self._points.shape
        if not self.n * self.d > 1:
            raise ValueError("`dataset` input should have multiple elements.")
        if self.d > 4:
            raise ValueError("the B200 evaluation kernel supports up to 4 dimensions")
        self._weights_dev = None
        if weights is not None:
            w = np.atleast_1d(np.asarray(weights.to_numpy() if isinstance(weights, ndarray) else weights)).astype(float)
            if w.ndim != 1:
                raise ValueError("`weights` input should be one-dimensional.")
            if len(w) != self.n:
                raise ValueError("`weights` input should be of length n")
            w = w / np.sum(w)
            self._weights = w.astype(np.float32)
            self._weights_dev = empty_on(ctx, (self.n,), np.float32)
            _lib.check(_lib.lib().cb_memcpy_h2d(ctx.handle, self._weights_dev.data_ptr, self._weights.ctypes.data,
                                                self._weights.nbytes))
            self._neff = 1.0 / float(np.sum(w ** 2))
        else:
            self._weights = None
            self._neff = float(self.n)
        self._covariance_factor = self._get_covariance_factor_function_from_bandwidth_type(bw_method)
        self._compute_covariance()
        self._replicas = {ctx.device: (self._whitened, self._weights_dev)}

    # ---- bandwidth and covariance (kde.py:1027-1081) --------------------------------------------------------
    @staticmethod
    def _get_covariance_factor_function_from_bandwidth_type(bw_method):
        if bw_method is None:
            return compute_scotts_factor
        if isinstance(bw_method, str):
            if bw_method == SCOTTS_FACTOR_STRING or bw_method == 'scott':
                return compute_scotts_factor
            if bw_method == SILVERMAN_FACTOR_STRING:
                return compute_silverman_factor
            raise ValueError(f'bw_method={bw_method} is not supported')
        if callable(bw_method):
            return bw_method
        if np.isscalar(bw_method):
            return lambda kde_info: bw_method
        raise ValueError(f'bw_method {bw_method} is not supported')

    def _compute_covariance(self):
        d = self.d
        raw = (ctypes.c_double * (1 + d + d * d))()
        _lib.check(_lib.lib().cb_kde_moments_f32(self._ctx.handle, self._points.data_ptr,
                                                 self._weights_dev.data_ptr if self._weights_dev is not None else None,
                                                 self.n, d, raw))
        raw = np.array(raw)
        if self._weights is None:                       # uniform weights 1/n
            wsum, sum_w2 = 1.0, 1.0 / self.n
            s1, s2 = raw[1:1 + d] / self.n, raw[1 + d:].reshape(d, d) / self.n
        else:
            wsum, sum_w2 = raw[0], float(np.sum(self._weights.astype(np.float64) ** 2))
            s1, s2 = raw[1:1 + d] / wsum, raw[1 + d:].reshape(d, d) / wsum
            sum_w2 /= wsum ** 2
        # numpy.cov(rowvar=True, bias=False, aweights=w): sum w (x-m)(x-m)^T / (1 - sum w^2) for normalised w
        self._data_covariance = np.atleast_2d((s2 - np.outer(s1, s1)) / (1.0 - sum_w2))
        self._data_inv_cov = np.linalg.inv(self._data_covariance)
        info = GaussianKDEInformation(dimension=d, n=self.n, neff=self.neff, points=self._points, weights=self._weights)
        self.factor = float(self._covariance_factor(info))
        self.covariance = self._data_covariance * self.factor ** 2
        self.inv_cov = self._data_inv_cov / self.factor ** 2
        self._norm_factor = math.sqrt(np.linalg.det(2 * np.pi * self.covariance))
        self._whitening = np.linalg.cholesky(self.inv_cov)                       # kde.py:1147-1154
        self._matrix = np.ascontiguousarray(self._whitening * _PRESCALE, dtype=np.float32)
        self._whitened = empty_on(self._ctx, (self.n, d), np.float32)            # kde.py:1156-1163
        _lib.check(_lib.lib().cb_kde_whiten_f32(self._ctx.handle, self._points.data_ptr, self.n, d,
                                                self._matrix.ctypes.data, self._whitened.data_ptr))

    # ---- properties of the reference ------------------------------------------------------------------------
    gpu = property(lambda self: self._gpu)
    neff = property(lambda self: self._neff)
    whitening = property(lambda self: self._whitening)
    whitened_points = property(lambda self: self._whitened)

    @property
    def weights(self) -> np.ndarray:
        return self._weights if self._weights is not None else np.full(self.n, 1.0 / self.n, dtype=np.float32)

    @property
    def dataset(self) -> ndarray:
        return self._points

    @property
    def normalization_constant(self) -> float:
        return (2 * np.pi) ** (-self.d / 2) * float(np.prod(np.diag(self._whitening)))

    # ---- evaluation -----------------------------------------------------------------------------------------
    def _replica(self, ctx):
        """Whitened data (and weights) on ``ctx``'s device; copied once per device."""
        if ctx.device not in self._replicas:
            wp = self._whitened.copy_to(ctx)                    # NVLink peer copy, once per device
            ww = self._weights_dev.copy_to(ctx) if self._weights_dev is not None else None
            self._replicas[ctx.device] = (wp, ww)
        return self._replicas[ctx.device]

    def _check_and_adjust_dimensions_of_points(self, points) -> np.ndarray:
        if isinstance(points, ndarray):
            points = points.to_numpy()
        points = np.atleast_2d(np.asarray(points))
        d, m = points.shape
        if d != self.d:
            if d == 1 and m == self.d:
                points = np.reshape(points, (self.d, 1))          # a single point passed as a row vector
            else:
                raise ValueError(f"points have dimension {d}, dataset has dimension {self.d}")
        return points

    def evaluate(self, points) -> ndarray:
        """Estimated pdf at ``points`` ((d, m) array, or a (d,) vector for one point): device array of m values."""
        points = self._check_and_adjust_dimensions_of_points(points)
        ctx = current_context()
        wp, ww = self._replica(ctx)
        m = points.shape[1]
        xi = np.ascontiguousarray(points.T.astype(np.float64) @ (self._whitening * _PRESCALE), dtype=np.float32)
        d_xi = empty_on(ctx, (m, self.d), np.float32)
        if m:
            _lib.check(_lib.lib().cb_memcpy_h2d(ctx.handle, d_xi.data_ptr, xi.ctypes.data, xi.nbytes))
        out = empty_on(ctx, (m,), np.float32)
        scale = self.normalization_constant * (1.0 if self._weights is not None else 1.0 / self.n)
        _lib.check(_lib.lib().cb_kde_evaluate_f32(ctx.handle, wp.data_ptr, ww.data_ptr if ww is not None else None,
                                                  self.n, self.d, d_xi.data_ptr, m, scale, out.data_ptr))
        return out

    __call__ = evaluate

    def pdf(self, x) -> ndarray:
        return self.evaluate(x)

    def logpdf(self, x) -> np.ndarray:
        return np.log(self.evaluate(x).to_numpy())

    # ---- integrals and resampling (kde.py:816-1025) --------------------------------------------------------
    def _gaussian_sums(self, sum_cov: np.ndarray, centres: np.ndarray) -> np.ndarray:
        """For every column c of ``centres`` (d, m): sum_i w_i exp(-(x_i - c)^T sum_cov^-1 (x_i - c) / 2) / norm, with
        norm = (2 pi)^(d/2) sqrt(det sum_cov) -- the dataset whitened with cholesky(sum_cov^-1) and pushed through the
        evaluation kernel (one MUFU.EX2 per (data point, centre) pair)."""
        ctx = self._ctx
        L = _lib.lib()
        d = self.d
        chol = np.linalg.cholesky(sum_cov)                               # raises LinAlgError if not s.p.d.
        norm_const = (2 * np.pi) ** (d / 2.0) * float(np.prod(np.diag(chol)))
        whitening = np.linalg.cholesky(np.linalg.inv(sum_cov))
        matrix = np.ascontiguousarray(whitening * _PRESCALE, dtype=np.float32)
        white = empty_on(ctx, (self.n, d), np.float32)
        _lib.check(L.cb_kde_whiten_f32(ctx.handle, self._points.data_ptr, self.n, d, matrix.ctypes.data, white.data_ptr))
        m = centres.shape[1]
        xi = np.ascontiguousarray(centres.T.astype(np.float64) @ (whitening * _PRESCALE), dtype=np.float32)
        d_xi = empty_on(ctx, (m, d), np.float32)
        _lib.check(L.cb_memcpy_h2d(ctx.handle, d_xi.data_ptr, xi.ctypes.data, xi.nbytes))
        out = empty_on(ctx, (m,), np.float32)
        scale = (1.0 if self._weights is not None else 1.0 / self.n) / norm_const
        _lib.check(L.cb_kde_evaluate_f32(ctx.handle, white.data_ptr,
                                         self._weights_dev.data_ptr if self._weights_dev is not None else None,
                                         self.n, d, d_xi.data_ptr, m, scale, out.data_ptr))
        return out.to_numpy().astype(np.float64)

    def integrate_gaussian(self, mean, cov) -> float:
        """Multiply the estimated density by a multivariate Gaussian and integrate over the whole space."""
        mean = np.atleast_1d(np.squeeze(np.asarray(mean, dtype=np.float64)))
        cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
        if mean.shape != (self.d,):
            raise ValueError("mean does not have dimension %s" % self.d)
        if cov.shape != (self.d, self.d):
            raise ValueError("covariance does not have dimension %s" % self.d)
        return float(self._gaussian_sums(self.covariance + cov, mean[:, None])[0])

    def integrate_box_1d(self, low, high) -> float:
        """Integral of a 1-D estimate between two bounds."""
        if self.d != 1:
            raise ValueError("integrate_box_1d() only handles 1D pdfs")
        stdev = float(np.sqrt(self.covariance).ravel()[0])
        value = ctypes.c_double(0.0)
        _lib.check(_lib.lib().cb_kde_integrate_box_1d_f32(
            self._ctx.handle, self._points.data_ptr,
            self._weights_dev.data_ptr if self._weights_dev is not None else None, self.n, float(low), float(high),
            stdev, ctypes.byref(value)))
        return value.value if self._weights is not None else value.value / self.n

    def integrate_kde(self, other: "gaussian_kde") -> float:
        """Integral of the product of this estimate with another one."""
        if other.d != self.d:
            raise ValueError("KDEs are not the same dimensionality")
        small, large = (other, self) if other.n < self.n else (self, other)
        centres = small._points.to_numpy().T.astype(np.float64)          # (d, n_small)
        sums = large._gaussian_sums(small.covariance + large.covariance, centres)
        return float(np.dot(sums, small.weights.astype(np.float64)))

    def resample(self, size: tp.Optional[int] = None, seed: tp.Optional[int] = None) -> ndarray:
        """Randomly sample a dataset from the estimated pdf: device array of shape (d, size).

        ``seed`` = None draws from the module generator (``cocos_b200.numerics.random``); an integer gives a
        reproducible draw without touching it (the reference accepts NumPy generators here; the device stream is
        Philox, so only integer seeds carry over)."""
        from ..numerics import random as _random
        if size is None:
            size = int(self.neff)
        size = int(size)
        if seed is None:
            seed_value, sub = _random.get_seed(), _random.next_subsequence()
        elif isinstance(seed, numbers.Integral):
            seed_value, sub = int(seed), 0
        else:
            raise TypeError("seed must be None or an integer (the device generator is Philox4x32-10)")
        ctx = self._ctx
        if self._weights is not None and getattr(self, "_cumweights_dev", None) is None:
            cum = np.cumsum(self._weights.astype(np.float64))
            self._cumweights_dev = empty_on(ctx, (self.n,), np.float64)
            _lib.check(_lib.lib().cb_memcpy_h2d(ctx.handle, self._cumweights_dev.data_ptr, cum.ctypes.data, cum.nbytes))
        chol = np.ascontiguousarray(np.linalg.cholesky(self.covariance), dtype=np.float32)
        out = empty_on(ctx, (self.d, size), np.float32)
        _lib.check(_lib.lib().cb_kde_resample_f32(
            ctx.handle, self._points.data_ptr,
            self._cumweights_dev.data_ptr if self._weights is not None else None, self.n, self.d, chol.ctypes.data,
            size, seed_value, sub, out.data_ptr))
        return out

    def evaluate_in_batches(self, points, maximum_number_of_elements_per_batch: int) -> np.ndarray:
        """Evaluate in batches of at most max/(n*d) points; results in main memory (kde.py:670-696)."""
        return evaluate_gaussian_kde_in_batches(self, points, maximum_number_of_elements_per_batch)

    def evaluate_in_batches_on_multiple_devices(self, points, maximum_number_of_elements_per_batch: int,
                                                compute_device_pool: ComputeDevicePool) -> np.ndarray:
        """Evaluation points split over the pool's GPUs with ``map_combine`` (kde.py:698-767)."""
        points = self._check_and_adjust_dimensions_of_points(points)
        m = points.shape[1]
        args_list = [[points[:, b:e]] for b, e in
                     generate_slices_with_number_of_batches(m, compute_device_pool.number_of_devices)] if m else []
        if not args_list:
            return np.zeros((0,), dtype=np.float32)

        def f(points_internal):
            return evaluate_gaussian_kde_in_batches(self, points_internal, maximum_number_of_elements_per_batch)
        return compute_device_pool.map_combine(f=f, combination=lambda parts: np.hstack(parts), args_list=args_list)


def _split_points_into_batches(points: np.ndarray, number_of_points_per_batch: int):
    m = points.shape[1]
    step = max(int(number_of_points_per_batch), 1)
    return [[points[:, b:min(b + step, m)]] for b in range(0, m, step)]


def evaluate_gaussian_kde_in_batches(kde: gaussian_kde, points, maximum_number_of_elements_per_batch: int) -> np.ndarray:
    points = kde._check_and_adjust_dimensions_of_points(points)
    per_batch = math.floor(maximum_number_of_elements_per_batch / (kde.n * kde.d))
    args_list = _split_points_into_batches(points, per_batch)
    if not args_list:
        return np.zeros((0,), dtype=np.float32)
    return map_combine_single_device(f=lambda p: kde.evaluate(p).to_numpy(),
                                     combination=lambda parts: np.hstack(parts), args_list=args_list)
