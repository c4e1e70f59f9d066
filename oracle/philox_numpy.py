"""TEST INFRASTRUCTURE ONLY -- never imported by the product (cocos_b200/).

CPU model (NumPy, vectorised) of the counter-based generator the CUDA kernels
use in place of ArrayFire's default engine.

The reference draws its GPU random numbers from ArrayFire's Philox 4x32-10
engine (cocos/numerics/random.py:30-51, cocos/options/__init__.py:10-16;
``af.data.randn`` / ``af.data.randu`` at random.py:137,152).  ArrayFire itself
is an un-vendored, un-pinned PyPI dependency (setup.py:24-34) that is absent
from /root/reference, and no reference test pins its output stream, so the
stream is "parity unpinned" at the ArrayFire boundary.  What IS pinned here is
the published Philox4x32-10 algorithm (Salmon et al., SC'11; constants as in
/usr/local/cuda/include/curand_philox4x32_x.h:88-91), checked against the
Random123 known-answer vectors in tests/golden/philox_kat.json.

Stream layout (shared with cocos_b200/csrc/philox.cuh, documented in DESIGN.md):
  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (index_lo, index_hi, block, subsequence)
  uniform  u01(o)  = asfloat(0x3f800000 | (o >> 9)) - 1      in [0,1)
  radius   U       = 2 - asfloat(0x3f800000 | (o >> 9))      in (0,1]
  angle    theta   = 2*pi*asfloat(0x3f800000 | (o' >> 9)) - 3*pi   in [-pi,pi)
  normals  z0 = sqrt(-2 ln U) cos(theta),  z1 = sqrt(-2 ln U) sin(theta)
The array kernels (rand/randn) draw two Box-Muller pairs per block from (o0,o1) and (o2,o3); the fused Heston
kernel splits one block into THREE steps: 23-bit radius from w_j >> 9, 19-bit angle from w_j[8:0]:w3 bits
(angle19 / box_muller_f64_from_fields below).
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32 with 10 rounds on arrays of counters; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for rnd in range(10):
        p0 = PHILOX_M0 * c0          # 32x32 -> 64 bit products, exact in uint64
        p1 = PHILOX_M1 * c2
        n0 = (p1 >> _SHIFT32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK32
        n2 = (p0 >> _SHIFT32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def philox_block(index, block, subsequence, seed):
    """Philox output for 64-bit ``index`` (array), 32-bit block and subsequence."""
    index = np.asarray(index, dtype=np.uint64)
    k0, k1 = split_seed(seed)
    return philox4x32_10(index & _MASK32, index >> _SHIFT32,
                         np.uint64(int(block) & 0xFFFFFFFF),
                         np.uint64(int(subsequence) & 0xFFFFFFFF), k0, k1)


def mantissa_float(o):
    """asfloat(0x3f800000 | (o >> 9)) in [1,2), float32, exact."""
    bits = (np.asarray(o, dtype=np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)
    return bits.view(np.float32)


def u01_f32(o):
    """Uniform in [0,1) with 23 random bits (float32, exact arithmetic)."""
    return mantissa_float(o) - np.float32(1.0)


def u01_f64(o_hi, o_lo):
    """Uniform in [0,1) with 52 random bits (float64, exact arithmetic)."""
    bits = (np.uint64(0x3FF0000000000000)
            | ((np.asarray(o_hi, dtype=np.uint64) & np.uint64(0xFFFFF)) << _SHIFT32)
            | np.asarray(o_lo, dtype=np.uint64))
    return bits.view(np.float64) - 1.0


def box_muller_f64(o_radius, o_angle):
    """Reference-precision (float64) evaluation of the kernels' Box-Muller map.

    The kernels evaluate the same map with MUFU.LG2/SQRT/SIN/COS in float32; the
    tests bound the difference.  Returns (z0, z1) = r*(cos, sin) in float64.
    """
    U = 2.0 - mantissa_float(o_radius).astype(np.float64)           # (0,1]
    t = mantissa_float(o_angle).astype(np.float64)                   # [1,2)
    theta = 2.0 * np.pi * t - 3.0 * np.pi                            # [-pi,pi)
    r = np.sqrt(-2.0 * np.log(U))
    return r * np.cos(theta), r * np.sin(theta)


def angle19(w_j, w3, j):
    """19-bit angle field of Euler step j (0..2) of a Philox block: w_j[8:0] : w3[31-10j : 22-10j]."""
    w_j = np.asarray(w_j, dtype=np.uint32)
    w3 = np.asarray(w3, dtype=np.uint32)
    return ((w_j & np.uint32(0x1FF)) << np.uint32(10)) | ((w3 >> np.uint32(22 - 10 * j)) & np.uint32(0x3FF))


def box_muller_f64_from_fields(radius23, angle19_field):
    """Box-Muller from a 23-bit radius mantissa and a 19-bit angle field (fused kernel, float64 model)."""
    t_r = ((np.asarray(radius23, dtype=np.uint32)) | np.uint32(0x3F800000)).view(np.float32)
    t_a = ((np.asarray(angle19_field, dtype=np.uint32) << np.uint32(4)) | np.uint32(0x3F800000)).view(np.float32)
    U = 2.0 - t_r.astype(np.float64)
    theta = 2.0 * np.pi * t_a.astype(np.float64) - 3.0 * np.pi
    r = np.sqrt(-2.0 * np.log(U))
    return r * np.cos(theta), r * np.sin(theta)


def box_muller_f64_from_u52(o0, o1, o2, o3):
    """float64 Box-Muller used by the f64 randn kernel: 52-bit uniforms."""
    U = 1.0 - u01_f64(o0, o1)                                        # (0,1]
    t = u01_f64(o2, o3)                                              # [0,1)
    theta = 2.0 * np.pi * t - np.pi
    r = np.sqrt(-2.0 * np.log(U))
    return r * np.cos(theta), r * np.sin(theta)


# --------------------------------------------------------------------------
# array-API models: what cb_rand_* / cb_randn_* write for element i
# --------------------------------------------------------------------------

def rand_f32(n, seed, subsequence):
    """Element i = u01(o[i % 4]) of Philox(index=i // 4, block=0)."""
    n = int(n)
    nblk = (n + 3) // 4
    o = philox_block(np.arange(nblk, dtype=np.uint64), 0, subsequence, seed)
    out = np.stack([u01_f32(x) for x in o], axis=1).reshape(-1)
    return out[:n]


def randn_f32_model(n, seed, subsequence):
    """float64 model of cb_randn_f32: elements 6j..6j+5 = the three (radius 23 bit, angle 19 bit) Box-Muller pairs
    of Philox block j (the split of the fused Heston kernel)."""
    n = int(n)
    nblk = (n + 5) // 6
    o = philox_block(np.arange(nblk, dtype=np.uint64), 0, subsequence, seed)
    cols = []
    for d in range(3):
        cols += list(box_muller_f64_from_fields(o[d] >> np.uint32(9), angle19(o[d], o[3], d)))
    return np.stack(cols, axis=1).reshape(-1)[:n]


def randn_pairs_f32_model(n, seed, subsequence):
    """float64 model of the two-pairs-per-block normals the antithetic kernels draw:
    elements 4j..4j+3 = (z0,z1)(o0,o1), (z0,z1)(o2,o3)."""
    n = int(n)
    nblk = (n + 3) // 4
    o = philox_block(np.arange(nblk, dtype=np.uint64), 0, subsequence, seed)
    a0, a1 = box_muller_f64(o[0], o[1])
    b0, b1 = box_muller_f64(o[2], o[3])
    return np.stack([a0, a1, b0, b1], axis=1).reshape(-1)[:n]


def rand_f64(n, seed, subsequence):
    """Element i = u01_f64(o[2*(i%2)], o[2*(i%2)+1]) of Philox(index=i // 2)."""
    n = int(n)
    nblk = (n + 1) // 2
    o = philox_block(np.arange(nblk, dtype=np.uint64), 0, subsequence, seed)
    out = np.stack([u01_f64(o[0], o[1]), u01_f64(o[2], o[3])], axis=1).reshape(-1)
    return out[:n]


def randn_f64_model(n, seed, subsequence):
    """Elements 2j, 2j+1 = Box-Muller pair from the four words of Philox(index=j)."""
    n = int(n)
    nblk = (n + 1) // 2
    o = philox_block(np.arange(nblk, dtype=np.uint64), 0, subsequence, seed)
    z0, z1 = box_muller_f64_from_u52(*o)
    return np.stack([z0, z1], axis=1).reshape(-1)[:n]


def pi_inside_count(n, seed, subsequence, antithetic=True):
    """Exact integer model of cb_pi_count.

    antithetic: draws H = ceil(n/2) points (ux,uy); point i<H is (ux_i,uy_i),
    point i+H (i < floor(n/2)) is (1-ux_i, 1-uy_i) -- the layout
    rand_antithetic([n], 0) produces (cocos/numerics/random.py:219-245) when
    x and y are drawn by two calls.  Draw j uses words (2*(j%2), 2*(j%2)+1) of
    Philox(index=j//2).  The test is rn(rn(x*x)+rn(y*y)) <= 1 in float32
    (examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py:28).
    """
    n = int(n)
    ndraw = (n + 1) // 2 if antithetic else n
    nblk = (ndraw + 1) // 2
    inside = 0
    chunk = 1 << 20
    for start in range(0, nblk, chunk):
        idx = np.arange(start, min(start + chunk, nblk), dtype=np.uint64)
        o = philox_block(idx, 0, subsequence, seed)
        ux = np.stack([u01_f32(o[0]), u01_f32(o[2])], axis=1).reshape(-1)
        uy = np.stack([u01_f32(o[1]), u01_f32(o[3])], axis=1).reshape(-1)
        j0 = 2 * start
        keep = ndraw - j0
        ux, uy = ux[:keep], uy[:keep]
        inside += int(np.count_nonzero(ux * ux + uy * uy <= np.float32(1.0)))
        if antithetic:
            nref = n // 2 - j0                      # partners exist for j < floor(n/2)
            ax = np.float32(1.0) - ux[:max(nref, 0)]
            ay = np.float32(1.0) - uy[:max(nref, 0)]
            inside += int(np.count_nonzero(ax * ax + ay * ay <= np.float32(1.0)))
    return inside
