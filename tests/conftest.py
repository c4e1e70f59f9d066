import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


def pytest_sessionstart(session):
    """(Re)build the CUDA library and the C oracle: both are git-ignored artefacts that nevertheless travel to the GPU
    box, so a stale binary could otherwise be tested against newer sources.  `make` is incremental (a no-op when the
    binaries are current); where the toolchain is absent (no nvcc / no make) an existing binary is used as it is."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "cocos_b200", "libcocos_b200.so")
    have_make = shutil.which("make") is not None
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if have_make and (os.path.exists(nvcc) or shutil.which("nvcc")):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "cocos_b200", "csrc")])
    elif not os.path.exists(lib):
        raise RuntimeError("libcocos_b200.so is missing and there is no nvcc/make to build it")
    if have_make and shutil.which("gcc"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    elif not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        raise RuntimeError("oracle/liboracle.so is missing and there is no gcc/make to build it")


@pytest.fixture(scope="session")
def golden_values():
    with open(os.path.join(GOLDEN, "reference_values.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_heston():
    return np.load(os.path.join(GOLDEN, "heston_injected_f32.npz"))


@pytest.fixture(scope="session")
def golden_antithetic():
    return np.load(os.path.join(GOLDEN, "antithetic_layouts.npz"))


@pytest.fixture(scope="session")
def ref_params(golden_values):
    return dict(golden_values["params"])


def heston_case(golden_heston, name, ref_params):
    """(R, N, Z, x, v, params) of one golden case."""
    R, N = (int(t) for t in golden_heston[name + "_RN"])
    if name == "alt":
        T, r, kappa, v_bar, sigma_v, rho, x0, v0 = (float(t) for t in golden_heston["alt_params"])
        p = dict(T=T, mu=r, kappa=kappa, v_bar=v_bar, sigma_v=sigma_v, rho=rho, x0=x0, v0=v0)
    else:
        q = ref_params
        p = dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
                 x0=q["x0"], v0=q["v0"])
    return R, N, golden_heston[name + "_Z"], golden_heston[name + "_x"], golden_heston[name + "_v"], p


HESTON_CASES = ("even", "odd", "two", "one", "long", "alt", "large")


def rel_err_in_price(x, x_ref):
    """max |exp(x) - exp(x_ref)| / exp(x_ref): the north_star's per-path terminal PRICE criterion."""
    a, b = np.exp(np.asarray(x, dtype=np.float64)), np.exp(np.asarray(x_ref, dtype=np.float64))
    return float(np.max(np.abs(a - b) / b))


@pytest.fixture(scope="session")
def gpu_device():
    from cocos_b200 import device
    device.init(device_id=0)
    return device
