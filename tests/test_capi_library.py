"""The C-ABI library: it loads, exports every symbol include/cocos_b200.h declares, fails loudly
without a GPU, and the product never routes through the oracle.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cocos_b200.h")).read()
    return sorted(set(re.findall(r"^CB_API\s+[\w\s\*]+?\b(cb_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 40
    for must in ("cb_heston_price_launch_f32", "cb_heston_price_launch_f64", "cb_heston_price_finish",
                 "cb_heston_terminal_injected_f32", "cb_heston_terminal_injected_f64", "cb_randn_antithetic_f32",
                 "cb_rand_antithetic_f32", "cb_pi_count_launch", "cb_comm_init_rank", "cb_allreduce_sum_f64",
                 "cb_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cocos_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert set(_lib.PROTOTYPES) == set(declared_symbols())       # the ctypes table covers the whole header
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    public = set(re.findall(r" T (\w+)", exported))
    assert public == set(declared_symbols()), public ^ set(declared_symbols())


def test_library_is_sm100a_only():
    from cocos_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    from cocos_b200 import _lib
    assert _lib.lib().cb_version().decode().startswith("cocos_b200")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.device_count()
    from cocos_b200 import heston
    from oracle.heston_numpy import REFERENCE_PARAMS
    with pytest.raises(RuntimeError):
        heston.simulate_and_compute_option_price(nT=11, R=100, **REFERENCE_PARAMS)
    import cocos_b200.numerics as cn
    with pytest.raises(RuntimeError):
        cn.random.randn(10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cocos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_product_does_not_import_torch():
    """libcocos_b200.so is the only native dependency of the product path: importing every product module (all but
    the torchrun helper cocos_b200.distributed, which needs torch.distributed for the rendezvous) leaves torch out."""
    import subprocess
    import sys
    code = ("import sys\n"
            "import cocos_b200, cocos_b200.heston, cocos_b200.monte_carlo_pi, cocos_b200.device\n"
            "import cocos_b200.numerics, cocos_b200.numerics.random, cocos_b200.numerics._array\n"
            "import cocos_b200.scientific.kde, cocos_b200.multi_processing.device_pool\n"
            "assert 'torch' not in sys.modules, 'torch was imported'\n"
            "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_plain_c_caller_links_against_the_abi(tmp_path):
    """examples/c/heston_price.c: a C program (no Python, no C++) compiles against include/cocos_b200.h and links
    against libcocos_b200.so; without a GPU it fails loudly with the library's own error message (no CPU fallback)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "heston_price"
    lib_dir = os.path.join(ROOT, "cocos_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-std=c11", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c", "heston_price.c"), "-L", lib_dir, "-lcocos_b200",
                           f"-Wl,-rpath,{lib_dir}", "-lm", "-o", str(exe)])
    out = subprocess.run([str(exe), "1000", "11"], capture_output=True, text=True, timeout=120)
    try:
        import ctypes
        have_gpu = ctypes.CDLL("libcuda.so.1").cuInit(0) == 0
    except OSError:
        have_gpu = False
    if have_gpu:
        assert out.returncode == 0 and "price" in out.stdout, out.stderr
    else:
        assert out.returncode != 0 and "no CPU fallback" in out.stderr
