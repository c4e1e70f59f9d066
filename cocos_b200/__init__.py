"""cocos_b200: the Heston Monte-Carlo hot path of michaelnowotny/cocos, rebuilt for NVIDIA B200.

Python host code over a ctypes C ABI (include/cocos_b200.h) to hand-written sm_100a
CUDA kernels; no ArrayFire, no multi-backend dispatch, no CPU fallback.

    import cocos_b200.numerics as cn                     # cn.random.randn / rand / randn_antithetic ...
    from cocos_b200.multi_processing import ComputeDevicePool
    from cocos_b200.heston import simulate_and_compute_option_price, simulate_and_compute_option_price_gpu
"""
__version__ = "0.1.0"

from . import device                                     # noqa: F401
from .numerics.numerical_package_bundle import B200Bundle, CocosBundle, NumericalPackageBundle  # noqa: F401
