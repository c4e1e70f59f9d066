"""cocos/numerics/numerical_package_selector.py:10-14 for the single B200 backend."""
from types import ModuleType


def select_num_pack(gpu: bool = True) -> ModuleType:
    if not gpu:
        raise NotImplementedError("cocos_b200 has no CPU branch (the NumPy restatement is test infrastructure in oracle/)")
    import cocos_b200.numerics as cn
    return cn
