"""Multi-GPU shard (needs >= 2 devices; skipped otherwise): disjoint pair ranges per GPU + ONE NCCL allreduce.

The correctness property is device-count invariance: every pair's Philox counter depends on its global index
only, so G GPUs reproduce the 1-GPU sums up to fp64 summation order."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        from cocos_b200 import _lib
        return _lib.device_count()
    except Exception:       # noqa: BLE001
        return 0


needs2 = pytest.mark.skipif(_device_count() < 2, reason="needs at least 2 GPUs")


@needs2
def test_pool_sums_are_device_count_invariant(gpu_device, ref_params):
    from cocos_b200 import _lib
    from cocos_b200.multi_processing import ComputeDevicePool
    q = ref_params
    p = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    R, N, strikes = 4_000_001, 101, [0.9, 0.95, 1.0]
    n = _device_count()
    one = ComputeDevicePool(compute_devices=[0])
    s1, q1 = one.heston_sums(p, R, N, strikes, seed=5, subsequence=3)
    for g in sorted({2, n}):
        pool = ComputeDevicePool(compute_devices=list(range(g)))
        assert pool.number_of_devices == g
        sg, qg = pool.heston_sums(p, R, N, strikes, seed=5, subsequence=3)
        np.testing.assert_allclose(sg, s1, rtol=1e-12)
        np.testing.assert_allclose(qg, q1, rtol=1e-12)
        s64, _ = pool.heston_sums(p, R // 4, N, strikes, seed=5, subsequence=3, dtype_suffix="f64")
        o64, _ = one.heston_sums(p, R // 4, N, strikes, seed=5, subsequence=3, dtype_suffix="f64")
        np.testing.assert_allclose(s64, o64, rtol=1e-12)
        pool.close()
    one.close()


@needs2
def test_reference_multi_gpu_entry_point(gpu_device, golden_values, ref_params):
    """simulate_and_compute_option_price_gpu with a pool (heston_utilities.py:217-270 call pattern)."""
    from cocos_b200 import heston
    from cocos_b200.multi_processing import ComputeDevicePool
    from cocos_b200.numerics import random as rnd
    pool = ComputeDevicePool()
    rnd.seed(11)
    price = heston.simulate_and_compute_option_price_gpu(nT=501, R=4_000_000, compute_device_pool=pool, **ref_params)
    ref = golden_values["statistical"]["nT501"]
    assert abs(price - ref["price"]) <= 3 * math.hypot(ref["se"], ref["se"] / 2)
    # generic map_reduce across devices: batch i on device i mod G, ordered fold
    devs = pool.map_reduce(f=lambda: __import__("cocos_b200").device.ComputeDeviceManager.get_current_compute_device_id(),
                           reduction=lambda acc, d: acc + [d], initial_value=[], number_of_batches=5)
    G = pool.number_of_devices
    assert devs == [i % G for i in range(5)]
    pool.close()


@needs2
def test_pi_sharded_over_pool(gpu_device):
    from cocos_b200 import monte_carlo_pi as mcp
    from cocos_b200.multi_processing import ComputeDevicePool
    from cocos_b200.numerics import random as rnd
    from oracle import c_oracle
    pool = ComputeDevicePool()
    n = 40_000_001
    for anti in (True, False):
        rnd.seed(21)
        got = mcp.count_inside(n, anti, compute_device_pool=pool)
        assert got == c_oracle.pi_inside(n, 21, 0, anti)         # bit-exact whatever the number of GPUs
    rnd.seed(21)
    assert abs(mcp.estimate_pi(400_000_000, batches=4, antithetic=True, compute_device_pool=pool) - math.pi) < 5e-4
    pool.close()


@needs2
def test_torchrun_bench_two_ranks(golden_values):
    """bench.py under torchrun: one rank per GPU, the cross-GPU sum inside the C library (fused or NCCL)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3",
           "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                  # rank 0 only
    res = json.loads(lines[0])
    assert res["n_gpus"] == 2 and res["config"]["paths"] == 20_000_000
    ref = golden_values["statistical"]["nT501"]
    assert abs(res["price"] - ref["price"]) <= 3 * math.hypot(res["price_se"], ref["se"])
    assert abs(res["e2e"]["price"] - ref["price"]) <= 3 * math.hypot(res["price_se"], ref["se"])
    assert res["config"]["collective"] in ("fused-peer-exchange", "nccl-allreduce")
    # the same job with the fused exchange switched off: same price to the last bits (two ranks: a+b either way)
    out2 = subprocess.run(cmd[:9] + ["29614"] + cmd[10:] + ["--no-cpu-baseline"], capture_output=True, text=True,
                          cwd=ROOT, env=dict(env, COCOS_B200_NO_P2P="1"), timeout=600)
    assert out2.returncode == 0, out2.stderr[-2000:]
    res2 = json.loads([l for l in out2.stdout.splitlines() if l.startswith("{")][0])
    assert res2["config"]["collective"] == "nccl-allreduce"
    assert abs(res2["price"] - res["price"]) <= 1e-12


def _multi(G, R, N, strikes, seed, sub, suffix="f32", spec=None):
    """cb_heston_price_multi_* called directly on freshly created contexts/communicators."""
    import ctypes
    from cocos_b200 import _lib
    L = _lib.lib()
    ctxs = [_lib.Context(d) for d in range(G)]
    harr = (ctypes.c_void_p * G)(*[c.handle for c in ctxs])
    comms = None
    if G > 1:
        comms = (ctypes.c_void_p * G)()
        _lib.check(L.cb_comm_init_all(G, harr, comms))
    k_arr, nK = _lib.strikes_array(strikes)
    s, q = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    p = _multi.params
    rc = getattr(L, f"cb_heston_price_multi_{suffix}")(G, harr, comms, ctypes.byref(p), R, N, k_arr, nK, seed, sub,
                                                       ctypes.byref(spec) if spec is not None else None, s, q)
    if comms is not None:
        for c in comms:
            L.cb_comm_destroy(ctypes.c_void_p(c))
    _lib.check(rc)
    return np.array(s), np.array(q)


@pytest.mark.parametrize("suffix", ["f32", "f64"])
def test_one_call_multi_entry_matches_launch_finish(gpu_device, ref_params, suffix):
    """The one-call C entry equals launch + finish on one device and is device-count invariant."""
    import ctypes
    from cocos_b200 import _lib
    q = ref_params
    _multi.params = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    R, N, strikes = 300_001, 64, [0.9, 0.95, 1.0, 1.05]
    s1, q1 = _multi(1, R, N, strikes, 9, 2, suffix)
    L = _lib.lib()
    ctx = _lib.Context(0)
    k_arr, nK = _lib.strikes_array(strikes)
    _lib.check(getattr(L, f"cb_heston_price_launch_{suffix}")(ctx.handle, ctypes.byref(_multi.params), R, N, k_arr, nK, 9, 2,
                                                              0, (R + 1) // 2, None, None, None, None))
    s0, q0 = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    _lib.check(L.cb_heston_price_finish(ctx.handle, nK, s0, q0))
    np.testing.assert_array_equal(s1, np.array(s0))
    np.testing.assert_array_equal(q1, np.array(q0))
    spec = _lib.PayoffSpec(1, 0, 0.0)          # Asian call through the same entry
    sa, _ = _multi(1, R, N, strikes, 9, 2, suffix, spec)
    assert np.all(sa > 0) and np.all(sa < s1)  # averaging lowers the call value under these parameters
    if _device_count() >= 2:
        s2, q2 = _multi(2, R, N, strikes, 9, 2, suffix)
        np.testing.assert_allclose(s2, s1, rtol=1e-12)
        np.testing.assert_allclose(q2, q1, rtol=1e-12)
        sa2, _ = _multi(2, R, N, strikes, 9, 2, suffix, spec)
        np.testing.assert_allclose(sa2, sa, rtol=1e-12)


def test_one_call_multi_entry_argument_errors(gpu_device, ref_params):
    import ctypes
    from cocos_b200 import _lib
    L = _lib.lib()
    q = ref_params
    p = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    ctx = _lib.Context(0)
    harr = (ctypes.c_void_p * 2)(ctx.handle, ctx.handle)
    k_arr, nK = _lib.strikes_array([1.0])
    s, sq = (ctypes.c_double * 1)(), (ctypes.c_double * 1)()
    assert L.cb_heston_price_multi_f32(0, harr, None, ctypes.byref(p), 100, 10, k_arr, nK, 0, 0, None, s, sq) != 0
    assert L.cb_heston_price_multi_f32(2, harr, None, ctypes.byref(p), 100, 10, k_arr, nK, 0, 0, None, s, sq) != 0
    assert b"communicator" in L.cb_last_error()
    inside = ctypes.c_uint64(0)
    assert L.cb_pi_count_multi(2, harr, None, 1000, 0, 0, 1, ctypes.byref(inside)) != 0
    _lib.check(L.cb_pi_count_multi(1, harr, None, 1000, 0, 0, 1, ctypes.byref(inside)))
    assert 700 < inside.value < 870


def test_nccl_loaded_before_torch_is_the_bundled_copy():
    """A communicator made before torch is imported must not pin an older system libnccl.so.2 under the SONAME
    torch's libtorch_cuda.so resolves against (regression: 'undefined symbol: ncclDevCommCreate')."""
    code = ("import ctypes, sys\n"
            "from cocos_b200 import _lib\n"
            "assert 'torch' not in sys.modules\n"
            "buf = ctypes.create_string_buffer(128)\n"
            "_lib.check(_lib.lib().cb_comm_unique_id(buf))\n"
            "import torch\n"
            "assert torch.cuda.is_available()\n"
            "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@needs2
def test_fused_peer_exchange_equals_nccl_allreduce(gpu_device, ref_params, monkeypatch):
    """The in-kernel cross-GPU sum over NVLink peer mappings and the NCCL allreduce behind the kernel give the same
    totals: several launches in a row (both slot parities), float and double, strike sweep, the conditional scheme,
    a path-dependent payoff and a job so small that one shard is empty."""
    from cocos_b200 import _lib
    from cocos_b200.multi_processing import ComputeDevicePool
    q = ref_params
    p = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    n = _device_count()

    def run(pool):
        out = []
        for sub in range(5):
            out.append(pool.heston_sums(p, 200_001, 33, [0.95], seed=3, subsequence=sub))
        out.append(pool.heston_sums(p, 100_000, 33, [0.8, 0.9, 1.0, 1.1], seed=3, subsequence=9))
        out.append(pool.heston_sums(p, 100_000, 33, [0.8, 0.9, 1.0, 1.1], seed=3, subsequence=9, dtype_suffix="f64"))
        out.append(pool.heston_sums(p, 100_000, 34, [0.95], seed=3, subsequence=10, scheme="conditional"))
        out.append(pool.heston_sums(p, 100_000, 33, [0.95], seed=3, subsequence=11, spec=_lib.PayoffSpec(1, 0, 0.0)))
        out.append(pool.heston_sums(p, 1, 33, [0.95], seed=3, subsequence=12))        # one pair: the other shards are empty
        out.append(pool.heston_sums(p, 200_001, 33, [0.95], seed=3, subsequence=13))  # and the exchange still lines up
        return out

    for g in sorted({2, n}):
        fused_pool = ComputeDevicePool(compute_devices=list(range(g)))
        if not fused_pool.peer_exchange_active:
            fused_pool.close()
            pytest.skip("no peer access between the devices of this box")
        fused = run(fused_pool)
        fused_pool.close()
        monkeypatch.setenv("COCOS_B200_NO_P2P", "1")
        nccl_pool = ComputeDevicePool(compute_devices=list(range(g)))
        assert not nccl_pool.peer_exchange_active
        via_nccl = run(nccl_pool)
        nccl_pool.close()
        monkeypatch.delenv("COCOS_B200_NO_P2P")
        for (s1, q1), (s2, q2) in zip(fused, via_nccl):
            np.testing.assert_allclose(s1, s2, rtol=1e-13)
            np.testing.assert_allclose(q1, q2, rtol=1e-13)


@needs2
def test_fused_exchange_detects_diverged_launch_sequences(gpu_device, ref_params):
    """Ranks whose calls differ do not get a silently wrong sum: the slots' tags (launch number, element count, call
    signature) are verified by the folding rank, and a mismatch surfaces as an error from cb_heston_price_finish."""
    import ctypes
    from cocos_b200 import _lib
    L = _lib.lib()
    q = ref_params
    p = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    ctxs = [_lib.Context(d) for d in range(2)]
    harr = (ctypes.c_void_p * 2)(*[c.handle for c in ctxs])
    comms = (ctypes.c_void_p * 2)()
    _lib.check(L.cb_comm_init_all(2, harr, comms))
    active = ctypes.c_int(0)
    _lib.check(L.cb_comm_p2p_active(ctypes.c_void_p(comms[0]), ctypes.byref(active)))
    if not active.value:
        for c in comms:
            L.cb_comm_destroy(ctypes.c_void_p(c))
        pytest.skip("no peer access between the devices of this box")
    k_arr, nK = _lib.strikes_array([0.95])
    R, N = 100_000, 17
    H = R // 2

    def launch(rank, sub):
        first = rank * (H // 2)
        return L.cb_heston_price_launch_f32(ctxs[rank].handle, ctypes.byref(p), R, N, k_arr, nK, 1, sub, first, H // 2,
                                            ctypes.c_void_p(comms[rank]), None, None, None)

    def finish(rank):
        s, sq = (ctypes.c_double * 1)(), (ctypes.c_double * 1)()
        rc = L.cb_heston_price_finish(ctxs[rank].handle, nK, s, sq)
        return rc, s[0]
    # matching calls: both ranks agree
    assert launch(0, 7) == 0 and launch(1, 7) == 0
    (rc0, s0), (rc1, s1) = finish(0), finish(1)
    assert rc0 == 0 and rc1 == 0 and s0 == s1 and s0 > 0
    # the ranks issue DIFFERENT calls under the same launch number (another subsequence): both must fail loudly
    assert launch(0, 8) == 0 and launch(1, 9) == 0
    (rc0, _), (rc1, _) = finish(0), finish(1)
    assert rc0 != 0 and rc1 != 0
    assert b"another launch" in L.cb_last_error()
    # and the communicator recovers: the next matching pair of calls is fine again
    assert launch(0, 10) == 0 and launch(1, 10) == 0
    (rc0, s0), (rc1, s1) = finish(0), finish(1)
    assert rc0 == 0 and rc1 == 0 and s0 == s1
    for c in comms:
        L.cb_comm_destroy(ctypes.c_void_p(c))


@needs2
def test_trajectory_and_pi_through_the_fused_exchange(gpu_device, ref_params):
    """Every launch type carries the cross-GPU sum inside the kernel: trajectory output and the pi count give the
    totals of the one-device run (no NCCL call is made once peer mappings exist)."""
    import ctypes
    from cocos_b200 import _lib
    from cocos_b200.numerics._array import empty_on
    from oracle import c_oracle
    L = _lib.lib()
    q = ref_params
    p = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    ctxs = [_lib.Context(d) for d in range(2)]
    harr = (ctypes.c_void_p * 2)(*[c.handle for c in ctxs])
    comms = (ctypes.c_void_p * 2)()
    _lib.check(L.cb_comm_init_all(2, harr, comms))
    k_arr, nK = _lib.strikes_array([0.95, 1.05])
    R, N = 40_000, 25
    H = R // 2
    # one device, whole job
    tx = empty_on(ctxs[0], (N, 2 * H), np.float32)
    x, v = empty_on(ctxs[0], (2 * H,), np.float32), empty_on(ctxs[0], (2 * H,), np.float32)
    _lib.check(L.cb_heston_trajectory_launch_f32(ctxs[0].handle, ctypes.byref(p), R, N, k_arr, nK, 2, 5, 0, H, None,
                                                 x.data_ptr, v.data_ptr, tx.data_ptr, None))
    s1, q1 = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    _lib.check(L.cb_heston_price_finish(ctxs[0].handle, nK, s1, q1))
    whole = tx.to_numpy()
    # two devices, half the pairs each, sums exchanged inside the kernels
    parts = []
    for r in range(2):
        t = empty_on(ctxs[r], (N, H), np.float32)
        xs, vs = empty_on(ctxs[r], (H,), np.float32), empty_on(ctxs[r], (H,), np.float32)
        _lib.check(L.cb_heston_trajectory_launch_f32(ctxs[r].handle, ctypes.byref(p), R, N, k_arr, nK, 2, 5, r * (H // 2),
                                                     H // 2, ctypes.c_void_p(comms[r]), xs.data_ptr, vs.data_ptr,
                                                     t.data_ptr, None))
        parts.append(t)
    for r in range(2):
        s2, q2 = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
        _lib.check(L.cb_heston_price_finish(ctxs[r].handle, nK, s2, q2))
        np.testing.assert_allclose(np.array(s2), np.array(s1), rtol=1e-12)
        np.testing.assert_allclose(np.array(q2), np.array(q1), rtol=1e-12)
        got = parts[r].to_numpy()
        lo = r * (H // 2)
        assert np.array_equal(got[:, :H // 2], whole[:, lo:lo + H // 2])               # first halves of the shard
        assert np.array_equal(got[:, H // 2:], whole[:, H + lo:H + lo + H // 2])       # their partners
    # pi: bit-exact count whatever the exchange path
    n = 30_000_001
    inside = ctypes.c_uint64(0)
    _lib.check(L.cb_pi_count_multi(2, harr, comms, n, 21, 4, 1, ctypes.byref(inside)))
    assert inside.value == c_oracle.pi_inside(n, 21, 4, True)
    for c in comms:
        L.cb_comm_destroy(ctypes.c_void_p(c))
