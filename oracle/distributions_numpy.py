"""TEST INFRASTRUCTURE ONLY -- NumPy model of cocos_b200/csrc/dist_kernels.cu.

Each function reproduces, from the same Philox words, the value the kernel computes for every element, using
float64 library functions where the kernel uses MUFU approximations (tests bound the difference), and the exact
float32 roundings where the kernel's arithmetic is exact (uniform, randint).  The transformations restate
cocos/numerics/random.py:252-531 (uniform :330-340, exponential :343-358, gamma (Marsaglia-Tsang) :372-425,
beta :431-441, wald :442-458, normal/lognormal :461-480, logistic :483-495, randint :252-287).
"""
import numpy as np

from . import philox_numpy as px


def _words4(n, seed, sub):
    nblk = (n + 3) // 4
    return px.philox_block(np.arange(nblk, dtype=np.uint64), 0, sub, seed)


def _u_open(w):
    return px.u01_f32(w).astype(np.float64) + 2.0 ** -24


def simple(kind, n, a, b, seed, sub):
    """kinds uniform / exponential / normal / lognormal / logistic: element 4j+e from word e of block j."""
    o = _words4(n, seed, sub)
    if kind in ("normal", "lognormal"):
        z = np.stack(px.box_muller_f64(o[0], o[1]) + px.box_muller_f64(o[2], o[3]), axis=1).reshape(-1)[:n]
        x = np.float32(a) + np.float32(b) * z
        return np.exp(x) if kind == "lognormal" else x
    w = np.stack(o, axis=1).reshape(-1)[:n]
    if kind == "uniform":
        return px.u01_f32(w) * np.float32(b - a) + np.float32(a)            # float32, bit-exact
    if kind == "exponential":
        return -float(np.float32(a)) * np.log((np.float32(1.0) - px.u01_f32(w)).astype(np.float64))
    if kind == "logistic":
        u = _u_open(w)
        return np.float32(a) + np.float64(np.float32(b)) * np.log(u / (1.0 - u))
    raise ValueError(kind)


def _gamma_mt(alpha, idx, stream, seed, sub):
    """Marsaglia-Tsang with attempts = Philox blocks, vectorised over the elements still rejecting."""
    d = alpha - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    v = np.ones(idx.shape)
    todo = np.ones(idx.shape, dtype=bool)
    for attempt in range(64):
        if not todo.any():
            break
        o = px.philox_block(idx[todo], attempt | (stream << 16), sub, seed)
        z, _ = px.box_muller_f64(o[0], o[1])
        y = 1.0 + c * z
        u = _u_open(o[2])
        vv = np.where(y > 0, y, 1.0) ** 3
        accept = (y > 0) & (np.log(u) < 0.5 * z * z + d - d * vv + d * np.log(vv))
        sel = np.flatnonzero(todo)
        # the kernel keeps the last positive-y candidate when all 64 attempts fail; never happens in practice
        v[sel[y > 0]] = vv[y > 0]
        todo[sel[accept]] = False
    return d * v


def gamma(n, shape, scale, seed, sub, stream=0):
    idx = np.arange(n, dtype=np.uint64)
    if shape >= 1.0:
        g = _gamma_mt(shape, idx, stream, seed, sub)
    else:
        o = px.philox_block(idx, stream << 16, sub, seed)
        g = _gamma_mt(shape + 1.0, idx, stream, seed, sub) * _u_open(o[3]) ** (1.0 / shape)
    return g * scale


def beta(n, a, b, seed, sub):
    x = gamma(n, a, 1.0, seed, sub, stream=0)
    y = gamma(n, b, 1.0, seed, sub, stream=1)
    return x / (x + y)


def wald(n, mean, scale, seed, sub):
    o = px.philox_block(np.arange(n, dtype=np.uint64), 0, sub, seed)
    z, _ = px.box_muller_f64(o[0], o[1])
    u = px.u01_f32(o[2]).astype(np.float64)
    y = z * z
    x = mean + mean ** 2 / (2.0 * scale) * y - mean / (2.0 * scale) * np.sqrt(4.0 * mean * scale * y + mean ** 2 * y * y)
    flip = u > mean / (mean + x)
    x[flip] = mean ** 2 / x[flip]
    return x


def randint(n, low, high, seed, sub):
    """trunc(u / f32(1/(high-low))) + low in float32 arithmetic: bit-exact model."""
    w = np.stack(_words4(n, seed, sub), axis=1).reshape(-1)[:n]
    divisor = np.float32(1.0 / (high - low))
    return (px.u01_f32(w) / divisor).astype(np.int64) + low
