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


def _source_hash(csrc):
    """sha256 over the sources the Makefile lists, in its order (the Makefile writes the same into libcocos_b200.srchash)."""
    import hashlib
    import re
    mk = open(os.path.join(csrc, "Makefile")).read()
    names = re.search(r"^SRCS\s*:=\s*(.*)$", mk, re.M).group(1).split() + \
        re.search(r"^HDRS\s*:=\s*(.*)$", mk, re.M).group(1).split() + ["Makefile"]
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(csrc, n), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def pytest_sessionstart(session):
    """Make sure the CUDA library and the C oracle are built FROM THE SOURCES AT HAND: both are git-ignored artefacts
    that nevertheless travel to the GPU box, so a stale binary could otherwise be tested against newer sources.
    The check is by content (a hash of the sources recorded at link time), not by mtime -- file times do not survive
    the copy to the GPU box.  Where the toolchain is absent an up-to-date binary is required."""
    import shutil
    import subprocess
    csrc = os.path.join(ROOT, "cocos_b200", "csrc")
    lib = os.path.join(ROOT, "cocos_b200", "libcocos_b200.so")
    stamp = os.path.join(ROOT, "cocos_b200", "libcocos_b200.srchash")
    current = os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == _source_hash(csrc)
    if not current:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if shutil.which("make") is None or not (os.path.exists(nvcc) or shutil.which("nvcc")):
            raise RuntimeError("libcocos_b200.so is missing or stale and there is no nvcc/make to rebuild it")
        subprocess.check_call(["make", "-s", "-B", "-j8", "-C", csrc])
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "heston_oracle.c")
    if not os.path.exists(so) or (os.path.getmtime(so) < os.path.getmtime(src) and shutil.which("gcc")):
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


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
