"""TEST INFRASTRUCTURE ONLY -- never imported by the product (cocos_b200/).

CPU restatement (NumPy) of the reference's Heston Monte-Carlo hot path.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker or the timed CPU baseline.

Pinned against the reference itself: tests/golden/*.npz were produced by running
the unmodified reference NumPy branch (oracle/make_golden.py, via
oracle/reference_runner.py) and tests/test_oracle_golden.py checks this
restatement (and oracle/heston_oracle.c) against them bit for bit.

What each function restates (paths relative to /root/reference):
  antithetic_slices            cocos/numerics/random.py:80-100
  check_shape_and_axis         cocos/numerics/random.py:103-120
  normal_antithetic            cocos/numerics/random.py:190-216
  uniform_antithetic           cocos/numerics/random.py:219-245
  heston_terminal              examples/stochastic_volatility/heston_utilities.py:58-96
  call_price                   examples/stochastic_volatility/heston_utilities.py:119
  price                        examples/stochastic_volatility/heston_utilities.py:122-164
  price_batched                heston_utilities.py:206-212, 260-268 (batch sizing, mean of batch prices)
  normalise_batches            cocos/multi_processing/utilities.py:88-106
  ordered_fold                 cocos/multi_processing/device_pool.py:246-251
  slices_with_batch_size / slices_with_number_of_batches
                               cocos/multi_processing/utilities.py:43-79
  estimate_pi                  examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py:18-33
"""
import math

import numpy as np

from . import philox_numpy as px

F32_EPS = np.finfo(np.float32).eps      # the floor is float32 eps even for float64 state (heston_utilities.py:91,94)

REFERENCE_PARAMS = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61,
                        kappa=6.21, v_bar=0.019, T=1.0, K=0.95)   # heston_pricing_benchmark.py:163-173


# --------------------------------------------------------------------------
# antithetic draws
# --------------------------------------------------------------------------

def antithetic_slices(shape, axis):
    """First floor(d/2) entries along ``axis``; everything along the other axes."""
    return tuple(slice(0, d // 2, 1) if k == axis else slice(0, d, 1)
                 for k, d in enumerate(shape))


def check_shape_and_axis(shape, axis):
    rank = len(shape)
    if rank > 4:
        raise ValueError('arrays with more than 4 axes are not supported')
    if rank < 1:
        raise ValueError('array must have at least one axis')
    if axis < 0 or axis > rank - 1:          # axis=None raises TypeError here, as the reference does
        raise ValueError(f'antithetic dimension must be None or between 0 and {rank}')


def _antithetic(shape, axis, dtype, draw, reflect):
    check_shape_and_axis(shape, axis)
    half = list(shape)
    half[axis] = math.ceil(shape[axis] / 2)
    first = draw(*half)
    if first.dtype != dtype:
        first = first.astype(dtype)
    mirrored = reflect(first[antithetic_slices(shape, axis)])
    return np.concatenate((first, mirrored), axis=axis)


def normal_antithetic(shape, axis, dtype=np.float32, draw=None):
    """(z, -z[:floor(n/2)]) along ``axis``; ``draw(*half_shape)`` supplies z (float64)."""
    return _antithetic(shape, axis, dtype, draw or np.random.randn, lambda z: -z)


def uniform_antithetic(shape, axis, dtype=np.float32, draw=None):
    """(u, 1-u[:floor(n/2)]) along ``axis``."""
    return _antithetic(shape, axis, dtype, draw or np.random.rand, lambda u: 1.0 - u)


# --------------------------------------------------------------------------
# Euler-Maruyama Heston paths (terminal values only, like the reference)
# --------------------------------------------------------------------------

def heston_terminal(T, N, R, mu, kappa, v_bar, sigma_v, rho, x0, v0,
                    dtype=np.float32, half_normals=None, record=None):
    """N-1 Euler steps of (log price x, variance v) for R antithetic paths.

    ``half_normals(step)`` -> float64 array (ceil(R/2), 2) of standard normals for
    step = 0..N-2 (None: numpy.random.randn, as the reference's NumPy branch).
    Operation order is the reference's, evaluated in ``dtype`` with Python-float
    scalars (weak scalars: they take the array dtype).  float32 is the reference's
    hard-coded dtype (heston_utilities.py:60-64); float64 is the C5 extension.
    """
    dt = T / float(N - 1)
    root_dt = math.sqrt(dt)
    mix = np.zeros((2,), dtype=dtype)
    mix[0] = rho
    mix[1] = math.sqrt(1 - rho ** 2)

    x = np.full((R,), x0, dtype=dtype)
    v = np.full((R,), v0, dtype=dtype)
    for step in range(N - 1):
        draw = (lambda *s: half_normals(step)) if half_normals is not None else None
        dB = normal_antithetic((R, 2), 0, dtype=dtype, draw=draw) * root_dt
        root_v = np.sqrt(v)
        x_next = x + (mu - 0.5 * v) * dt + np.multiply(root_v, dB[:, 0])
        v_next = v + kappa * (v_bar - v) * dt + sigma_v * np.multiply(root_v, np.dot(dB, mix))
        v = np.maximum(v_next, F32_EPS)
        x = x_next
        if record is not None:
            record.append(x.copy())         # log prices at the simulated dates 1..N-1 (path-dependent payoffs)
    return x, np.maximum(v, F32_EPS)


def call_price(r, T, K, x_T):
    """Discounted mean call payoff, the reference's expression (array dtype mean)."""
    return float(math.exp(-r * T) * np.mean(np.maximum(np.exp(x_T) - K, 0)))


def call_price_and_se(r, T, K, x_T):
    """Price in float64 plus the standard error from antithetic PAIR averages."""
    R = x_T.shape[0]
    H = (R + 1) // 2
    pay = np.maximum(np.exp(x_T.astype(np.float64)) - K, 0.0)
    unit = pay[:H].copy()
    unit[:R // 2] = 0.5 * (pay[:R // 2] + pay[H:])
    disc = math.exp(-r * T)
    return disc * pay.mean(), disc * unit.std(ddof=1) / math.sqrt(H)


def price(x0, v0, r, rho, sigma_v, kappa, v_bar, T, K, nT, R, dtype=np.float32, half_normals=None):
    x_T, _ = heston_terminal(T=T, N=nT, R=R, mu=r, kappa=kappa, v_bar=v_bar, sigma_v=sigma_v,
                             rho=rho, x0=x0, v0=v0, dtype=dtype, half_normals=half_normals)
    return call_price(r, T, K, x_T)


# --------------------------------------------------------------------------
# map-reduce semantics
# --------------------------------------------------------------------------

def normalise_batches(args_list=None, kwargs_list=None, number_of_batches=None):
    if number_of_batches is None:
        if args_list is not None:
            number_of_batches = len(args_list)
        elif kwargs_list is not None:
            number_of_batches = len(kwargs_list)
        else:
            raise ValueError('Number_of_batches must be defined if '
                             'both args_list and kwargs_list are empty')
    if args_list is None:
        args_list = number_of_batches * [list()]
    if kwargs_list is None:
        kwargs_list = number_of_batches * [dict()]
    return args_list, kwargs_list, number_of_batches


def ordered_fold(indexed_results, reduction, initial_value):
    """Left fold in submission order, whatever order the results completed in."""
    acc = initial_value
    for _, value in sorted(indexed_results, key=lambda pair: pair[0]):
        acc = reduction(acc, value)
    return acc


def price_batched(number_of_batches, R, **kw):
    """Mean of ``number_of_batches`` batch prices of ceil(R/nb) paths each."""
    per_batch = math.ceil(R / number_of_batches)
    results = [(i, price(R=per_batch, **kw)) for i in range(number_of_batches)]
    return ordered_fold(results, lambda acc, p: acc + p / number_of_batches, 0.0)


def slices_with_batch_size(n, batch_size):
    out, lo = [], 0
    while lo < n:
        out.append((lo, min(lo + batch_size, n)))
        lo += batch_size
    return tuple(out)


def slices_with_number_of_batches(n, number_of_batches):
    return slices_with_batch_size(n, math.ceil(n / number_of_batches))


# --------------------------------------------------------------------------
# Monte-Carlo pi
# --------------------------------------------------------------------------

def estimate_pi(n, batches=1, draw=None):
    """4 * mean(x^2 + y^2 <= 1) per batch of ceil(n/batches) float32 points, averaged."""
    draw = draw or np.random.rand
    per_batch = math.ceil(n / batches)
    acc = 0.0
    for _ in range(batches):
        x = draw(per_batch).astype(np.float32)
        y = draw(per_batch).astype(np.float32)
        acc += 4.0 * float(np.mean((x * x + y * y) <= 1.0))
    return acc / batches


# --------------------------------------------------------------------------
# model of the fused kernel: Philox stream -> Box-Muller (float64) -> paths
# --------------------------------------------------------------------------

def philox_half_normals(R, seed, subsequence, pair_offset=0):
    """Returns half_normals(step) reproducing the fused kernel's draws.

    One Philox block feeds three Euler steps (cocos_b200/csrc/heston_fused.cu: split_draws).  Pair p (global
    index pair_offset + p) at Euler step t (0-based) uses block t//3 of Philox(index=pair, block, subsequence);
    with j = t % 3 and words (w0, w1, w2, w3):
      radius mantissa (23 bits) = w_j >> 9
      angle  mantissa (19 bits) = (w_j & 0x1ff) << 10 | (w3 >> (22 - 10 j)) & 0x3ff, placed at mantissa bits [22:4]
    """
    H = (R + 1) // 2
    idx = np.arange(H, dtype=np.uint64) + np.uint64(pair_offset)
    cache = {}

    def half_normals(step):
        blk, j = divmod(step, 3)
        if blk not in cache:
            cache.clear()
            cache[blk] = px.philox_block(idx, blk, subsequence, seed)
        w = cache[blk]
        z0, z1 = px.box_muller_f64_from_fields(w[j] >> np.uint32(9), px.angle19(w[j], w[3], j))
        return np.stack([z0, z1], axis=1)
    return half_normals


def fused_model(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, R, seed, subsequence,
                pair_offset=0, dtype=np.float32):
    """Terminal (x, v) the fused kernel should reproduce up to float rounding."""
    return heston_terminal(T=T, N=nT, R=R, mu=r, kappa=kappa, v_bar=v_bar, sigma_v=sigma_v,
                           rho=rho, x0=x0, v0=v0, dtype=dtype,
                           half_normals=philox_half_normals(R, seed, subsequence, pair_offset))


def payoff_sums(x_T, strikes):
    """(sum of payoffs, sum of squared pair-average payoffs) per strike, float64.

    The last draw of an odd R has no partner: its unit is the path itself.
    """
    R = x_T.shape[0]
    H = (R + 1) // 2
    S = np.exp(x_T.astype(np.float64))
    sums, sumsq = [], []
    for K in np.atleast_1d(np.asarray(strikes, dtype=np.float64)):
        pay = np.maximum(S - K, 0.0)
        unit = pay[:H].copy()
        unit[:R // 2] = 0.5 * (pay[:R // 2] + pay[H:])
        sums.append(pay.sum())
        sumsq.append((unit * unit).sum())
    return np.array(sums), np.array(sumsq)


# --------------------------------------------------------------------------
# model of the conditional-Gaussian kernel (cocos_b200/csrc/heston_conditional.cu)
# --------------------------------------------------------------------------

def conditional_model(x0, v0, r, rho, sigma_v, kappa, v_bar, T, nT, R, seed, subsequence, pair_offset=0):
    """Terminal (x, v) of the opt-in conditional scheme from the same Philox words, in float64.

    Law-equivalent to heston_terminal (the Euler scheme of heston_utilities.py:79-91): the variance recursion is
    driven by w = rho z0 + rho' z1 only, the orthogonal Gaussian part of each pair is drawn once at the end from
    block 0xFFFFFFFF.  One Philox block feeds six steps: draw j (0..2) gives (r cos, r sin) for steps 6b+2j, 6b+2j+1.
    """
    steps = nT - 1
    dt = T / steps
    H, F = (R + 1) // 2, R // 2
    idx = np.arange(H, dtype=np.uint64) + np.uint64(pair_offset)
    eps = float(F32_EPS)
    v1 = np.full(H, v0, dtype=np.float64)
    v2 = v1.copy()
    A = np.zeros(H); B = np.zeros(H); C = np.zeros(H); P1 = np.zeros(H); P2 = np.zeros(H)
    a, b = 1.0 - kappa * dt, kappa * v_bar * dt
    scale = sigma_v * math.sqrt(dt)
    t = 0
    blk = 0
    while t < steps:
        w = px.philox_block(idx, blk, subsequence, seed)
        for j in range(3):
            za, zb = px.box_muller_f64_from_fields(w[j] >> np.uint32(9), px.angle19(w[j], w[3], j))
            for z in (za, zb):
                if t >= steps:
                    break
                ws = scale * z
                s1, s2 = np.sqrt(v1), np.sqrt(v2)
                A += v1; B += v2; C += s1 * s2; P1 += s1 * ws; P2 += s2 * ws
                v1 = np.maximum(a * v1 + b + s1 * ws, eps)
                v2 = np.maximum(a * v2 + b - s2 * ws, eps)
                t += 1
        blk += 1
    fin = px.philox_block(idx, 0xFFFFFFFF, subsequence, seed)
    z1, z2 = px.box_muller_f64(fin[0], fin[1])
    rootA = np.sqrt(A)
    G1 = rootA * z1
    cross = np.where(A > 0, C / np.where(A > 0, rootA, 1.0), 0.0)
    G2 = -cross * z1 + np.sqrt(np.maximum(B - cross * cross, 0.0)) * z2
    base = x0 + steps * (r * dt)
    orth = math.sqrt(max(1.0 - rho * rho, 0.0)) * math.sqrt(dt)
    xa = base - 0.5 * dt * A + (rho / sigma_v) * P1 + orth * G1
    xb = base - 0.5 * dt * B - (rho / sigma_v) * P2 + orth * G2
    return np.concatenate([xa, xb[:F]]), np.concatenate([v1, v2[:F]])


def path_dependent_sums(x_path, strikes, kind="asian", put=False, barrier=None):
    """(sum of payoffs, sum of squared pair-unit averages) for path-dependent payoffs on recorded log-price paths.

    x_path: (N-1, R) log prices at the simulated dates (heston_terminal(record=[...])).  Textbook definitions:
    asian = arithmetic average of the N-1 prices; up/down-and-out = knocked out when any monitored price is
    >= / <= the barrier, otherwise the European payoff.
    """
    X = np.asarray(x_path, dtype=np.float64)
    S = np.exp(X)
    R = X.shape[1]
    H = (R + 1) // 2
    terminal = S[-1]
    if kind == "asian":
        level, alive = S.mean(axis=0), np.ones(R, dtype=bool)
    elif kind == "up_and_out":
        level, alive = terminal, X.max(axis=0) < math.log(barrier)
    elif kind == "down_and_out":
        level, alive = terminal, X.min(axis=0) > math.log(barrier)
    else:
        level, alive = terminal, np.ones(R, dtype=bool)
    sums, sumsq = [], []
    for K in np.atleast_1d(np.asarray(strikes, dtype=np.float64)):
        pay = np.maximum((K - level) if put else (level - K), 0.0) * alive
        unit = pay[:H].copy()
        unit[:R // 2] = 0.5 * (pay[:R // 2] + pay[H:])
        sums.append(pay.sum())
        sumsq.append((unit * unit).sum())
    return np.array(sums), np.array(sumsq)
