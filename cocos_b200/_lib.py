"""ctypes binding of libcocos_b200.so (C ABI: include/cocos_b200.h).

There is no CPU fallback: if the shared library has not been built the import of
any compute function fails with ImportError, and without a CUDA device every
compute call raises RuntimeError (CB_ERR_CUDA).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# $COCOS_B200_LIBRARY names another build of the same sources (A/B runs of tuning variants); default: the in-tree one
LIB_PATH = os.environ.get("COCOS_B200_LIBRARY") or os.path.join(_HERE, "libcocos_b200.so")

c_int, c_u32, c_u64, c_size = ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
c_void_p, c_char_p, c_double, c_float = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double, ctypes.c_float
P = ctypes.POINTER

MAX_STRIKES = 64
UNIQUE_ID_BYTES = 128
P2P_HANDLE_BYTES = 64

CB_OK, CB_ERR_INVALID, CB_ERR_CUDA, CB_ERR_NCCL, CB_ERR_UNSUPPORTED = range(5)


class HestonParams(ctypes.Structure):
    """cb_heston_params: the scalars of simulate_heston_model (heston_utilities.py:24-35)."""
    _fields_ = [(n, c_double) for n in ("T", "mu", "kappa", "v_bar", "sigma_v", "rho", "x0", "v0")]


class PayoffSpec(ctypes.Structure):
    """cb_payoff_spec: kind 0 european, 1 asian, 2 up-and-out, 3 down-and-out."""
    _fields_ = [("kind", c_int), ("is_put", c_int), ("barrier", c_double)]


# name -> (restype, argtypes); every symbol include/cocos_b200.h declares
PROTOTYPES = {
    "cb_last_error": (c_char_p, []),
    "cb_version": (c_char_p, []),
    "cb_device_count": (c_int, [P(c_int)]),
    "cb_device_name": (c_int, [c_int, c_char_p, c_size]),
    "cb_device_attributes": (c_int, [c_int, P(c_int), P(c_int), P(c_int), P(c_size)]),
    "cb_ctx_create": (c_int, [c_int, c_void_p, P(c_void_p)]),
    "cb_ctx_destroy": (c_int, [c_void_p]),
    "cb_ctx_sync": (c_int, [c_void_p]),
    "cb_ctx_device": (c_int, [c_void_p, P(c_int)]),
    "cb_ctx_stream": (c_void_p, [c_void_p]),
    "cb_timer_start": (c_int, [c_void_p]),
    "cb_timer_stop": (c_int, [c_void_p, P(c_float)]),
    "cb_ctx_set_launch": (c_int, [c_void_p, c_int, c_int, c_int]),
    "cb_ctx_set_variant": (c_int, [c_void_p, c_int]),
    "cb_malloc": (c_int, [c_void_p, c_size, P(c_void_p)]),
    "cb_free": (c_int, [c_void_p, c_void_p]),
    "cb_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size]),
    "cb_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size]),
    "cb_memcpy_d2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size]),
    "cb_memcpy_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size]),
    "cb_cast": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_u64]),
    "cb_unary": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_int, c_int]),
    "cb_transpose": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_u64, c_int]),
    "cb_philox_words": (c_int, [c_void_p, c_void_p, c_u64, c_u64, c_u32, c_u64, c_u32]),
    "cb_rand_f32": (c_int, [c_void_p, c_void_p, c_u64, c_u64, c_u32]),
    "cb_rand_f64": (c_int, [c_void_p, c_void_p, c_u64, c_u64, c_u32]),
    "cb_randn_f32": (c_int, [c_void_p, c_void_p, c_u64, c_u64, c_u32]),
    "cb_randn_f64": (c_int, [c_void_p, c_void_p, c_u64, c_u64, c_u32]),
    "cb_randn_antithetic_f32": (c_int, [c_void_p, c_void_p, P(c_u64), c_int, c_int, c_u64, c_u32]),
    "cb_randn_antithetic_f64": (c_int, [c_void_p, c_void_p, P(c_u64), c_int, c_int, c_u64, c_u32]),
    "cb_rand_antithetic_f32": (c_int, [c_void_p, c_void_p, P(c_u64), c_int, c_int, c_u64, c_u32]),
    "cb_rand_antithetic_f64": (c_int, [c_void_p, c_void_p, P(c_u64), c_int, c_int, c_u64, c_u32]),
    "cb_random_distribution_f32": (c_int, [c_void_p, c_void_p, c_u64, c_int, c_double, c_double, c_u64, c_u32]),
    "cb_randint": (c_int, [c_void_p, c_void_p, c_int, c_u64, ctypes.c_int64, ctypes.c_int64, c_u64, c_u32]),
    "cb_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_u64, c_int]),
    "cb_multivariate_normal_f32": (c_int, [c_void_p, c_void_p, c_u64, c_int, c_void_p, c_void_p, c_u64, c_u32]),
    "cb_multivariate_normal_f64": (c_int, [c_void_p, c_void_p, c_u64, c_int, c_void_p, c_void_p, c_u64, c_u32]),
    "cb_kde_moments_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_int, c_void_p]),
    "cb_kde_whiten_f32": (c_int, [c_void_p, c_void_p, c_u64, c_int, c_void_p, c_void_p]),
    "cb_kde_evaluate_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_int, c_void_p, c_u64, c_double, c_void_p]),
    "cb_kde_integrate_box_1d_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_double, c_double, c_double, P(c_double)]),
    "cb_kde_resample_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_int, c_void_p, c_u64, c_u64, c_u32, c_void_p]),
    "cb_heston_terminal_injected_f32": (c_int, [c_void_p, c_void_p, P(HestonParams), c_u64, c_u32, c_void_p, c_void_p]),
    "cb_heston_terminal_injected_f64": (c_int, [c_void_p, c_void_p, P(HestonParams), c_u64, c_u32, c_void_p, c_void_p]),
    "cb_heston_terminal_injected_host_f32": (c_int, [c_void_p, c_void_p, P(HestonParams), c_u64, c_u32, c_void_p, c_void_p]),
    "cb_heston_price_launch_f32": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64, c_u32,
                                           c_u64, c_u64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_heston_price_launch_f64": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64, c_u32,
                                           c_u64, c_u64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_heston_trajectory_launch_f32": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64, c_u32,
                                                c_u64, c_u64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_heston_trajectory_launch_f64": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64, c_u32,
                                                c_u64, c_u64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_heston_price_conditional_launch_f32": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64,
                                                       c_u32, c_u64, c_u64, c_void_p, c_void_p, c_void_p]),
    "cb_heston_payoff_launch_f32": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64, c_u32,
                                            c_u64, c_u64, c_void_p, P(PayoffSpec)]),
    "cb_heston_payoff_launch_f64": (c_int, [c_void_p, P(HestonParams), c_u64, c_u32, P(c_double), c_int, c_u64, c_u32,
                                            c_u64, c_u64, c_void_p, P(PayoffSpec)]),
    "cb_heston_price_finish": (c_int, [c_void_p, c_int, P(c_double), P(c_double)]),
    "cb_heston_price_multi_f32": (c_int, [c_int, P(c_void_p), P(c_void_p), P(HestonParams), c_u64, c_u32, P(c_double),
                                          c_int, c_u64, c_u32, P(PayoffSpec), P(c_double), P(c_double)]),
    "cb_heston_price_multi_f64": (c_int, [c_int, P(c_void_p), P(c_void_p), P(HestonParams), c_u64, c_u32, P(c_double),
                                          c_int, c_u64, c_u32, P(PayoffSpec), P(c_double), P(c_double)]),
    "cb_call_payoff_sums_f32": (c_int, [c_void_p, c_void_p, c_u64, P(c_double), c_int, P(c_double), P(c_double)]),
    "cb_call_payoff_sums_f64": (c_int, [c_void_p, c_void_p, c_u64, P(c_double), c_int, P(c_double), P(c_double)]),
    "cb_pi_count_launch": (c_int, [c_void_p, c_u64, c_u64, c_u32, c_int, c_u64, c_u64, c_void_p]),
    "cb_pi_count_finish": (c_int, [c_void_p, P(c_u64)]),
    "cb_pi_count_multi": (c_int, [c_int, P(c_void_p), P(c_void_p), c_u64, c_u64, c_u32, c_int, P(c_u64)]),
    "cb_nccl_library": (c_int, [ctypes.c_char_p]),
    "cb_comm_unique_id": (c_int, [c_void_p]),
    "cb_comm_init_rank": (c_int, [c_void_p, c_int, c_int, c_void_p, P(c_void_p)]),
    "cb_comm_init_all": (c_int, [c_int, P(c_void_p), P(c_void_p)]),
    "cb_comm_destroy": (c_int, [c_void_p]),
    "cb_comm_p2p_export": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cb_comm_p2p_import": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cb_comm_p2p_enable": (c_int, [c_void_p, c_int]),
    "cb_comm_p2p_active": (c_int, [c_void_p, P(c_int)]),
    "cb_comm_p2p_disconnect": (c_int, [c_void_p]),
    "cb_allreduce_sum_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_size]),
    "cb_group_start": (c_int, []),
    "cb_group_end": (c_int, []),
    "cb_peak_measure": (c_int, [c_void_p, c_int, c_int, P(c_double), P(c_float)]),
}

_lib = None
_lib_lock = threading.Lock()


def lib():
    """The loaded shared library (loads on first use, no CUDA call is made)."""
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make -C cocos_b200/csrc`). cocos_b200 has no CPU fallback.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (restype, argtypes) in PROTOTYPES.items():
                    fn = getattr(handle, name)      # AttributeError if the header and the library disagree
                    fn.restype = restype
                    fn.argtypes = argtypes
                bundled = _bundled_nccl()
                if bundled:
                    handle.cb_nccl_library(bundled.encode())
                _lib = handle
    return _lib


def _bundled_nccl():
    """Path of the libnccl.so.2 torch ships (found without importing torch): torch's libtorch_cuda.so needs that
    version, so it must be the copy this process maps first."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        return None
    for root in (spec.submodule_search_locations if spec else []):
        path = os.path.join(root, "lib", "libnccl.so.2")
        if os.path.exists(path):
            return path
    return None


_ERRORS = {CB_ERR_INVALID: ValueError, CB_ERR_CUDA: RuntimeError, CB_ERR_NCCL: RuntimeError,
           CB_ERR_UNSUPPORTED: NotImplementedError}


def check(rc):
    """Raise the Python exception the reference would raise for the same condition."""
    if rc != CB_OK:
        msg = lib().cb_last_error().decode("utf-8", "replace")
        raise _ERRORS.get(rc, RuntimeError)(msg)


def device_count():
    n = c_int(0)
    check(lib().cb_device_count(ctypes.byref(n)))
    return n.value


class Context:
    """Python owner of a cb_ctx (one device, one stream, reduction workspace)."""

    def __init__(self, device, stream=None):
        self.device = int(device)
        handle = c_void_p()
        check(lib().cb_ctx_create(self.device, c_void_p(stream) if stream else None, ctypes.byref(handle)))
        self.handle = handle

    @property
    def stream_ptr(self):
        """The context's cudaStream_t as an integer (for frameworks that want to enqueue work behind the kernels)."""
        return lib().cb_ctx_stream(self.handle)

    def sync(self):
        check(lib().cb_ctx_sync(self.handle))

    def set_launch(self, threads=0, blocks_per_sm=0, pairs_per_thread=0):
        check(lib().cb_ctx_set_launch(self.handle, threads, blocks_per_sm, pairs_per_thread))

    def set_variant(self, variant=0):
        check(lib().cb_ctx_set_variant(self.handle, variant))

    def timer_start(self):
        check(lib().cb_timer_start(self.handle))

    def timer_stop(self):
        ms = c_float(0)
        check(lib().cb_timer_stop(self.handle, ctypes.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "handle", None):
            lib().cb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}
_ctx_lock = threading.Lock()


def context(device):
    """Process-wide context of a device (created on first use)."""
    device = int(device)
    ctx = _contexts.get(device)
    if ctx is None:
        with _ctx_lock:
            ctx = _contexts.get(device)
            if ctx is None:
                ctx = _contexts[device] = Context(device)
    return ctx


def strikes_array(strikes):
    vals = [float(k) for k in strikes]
    if not 1 <= len(vals) <= MAX_STRIKES:
        raise ValueError(f"number of strikes must be in [1,{MAX_STRIKES}]")
    return (c_double * len(vals))(*vals), len(vals)
