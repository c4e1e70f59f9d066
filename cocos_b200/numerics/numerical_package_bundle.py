"""The drop-in boundary: cocos/numerics/numerical_package_bundle.py:6-29 protocol.

The reference selects its backend by passing a ``NumericalPackageBundle`` subclass
into the Heston functions (heston_utilities.py:35,104,134).  ``B200Bundle`` is the
bundle of this package; ``CocosBundle`` is kept as an alias so that call sites
written for the reference pick up the B200 path unchanged.
"""
from abc import ABC, abstractmethod
from types import ModuleType
import typing as tp


class NumericalPackageBundle(ABC):
    @classmethod
    @abstractmethod
    def is_installed(cls) -> bool:
        ...

    @classmethod
    @abstractmethod
    def label(cls) -> str:
        ...

    @classmethod
    @abstractmethod
    def module(cls) -> ModuleType:
        ...

    @classmethod
    @abstractmethod
    def random_module(cls) -> ModuleType:
        ...

    @classmethod
    def synchronize(cls):
        pass


class B200Bundle(NumericalPackageBundle):
    @classmethod
    def is_installed(cls) -> bool:
        import os
        from .. import _lib
        return os.path.exists(_lib.LIB_PATH)

    @classmethod
    def label(cls) -> str:
        return 'Cocos-B200'

    @classmethod
    def module(cls) -> ModuleType:
        import cocos_b200.numerics
        return cocos_b200.numerics

    @classmethod
    def random_module(cls) -> ModuleType:
        import cocos_b200.numerics.random
        return cocos_b200.numerics.random

    @classmethod
    def synchronize(cls):
        from ..device import sync
        sync()


CocosBundle = B200Bundle


class NumpyBundle(NumericalPackageBundle):
    """numerical_package_bundle.py:60-80 of the reference, kept so that code enumerating bundles still imports.
    The Heston entry points of this package refuse it: there is no CPU branch here (see oracle/ for the restatement)."""

    @classmethod
    def is_installed(cls) -> bool:
        return True

    @classmethod
    def label(cls) -> str:
        return 'NumPy'

    @classmethod
    def module(cls) -> ModuleType:
        import numpy
        return numpy

    @classmethod
    def random_module(cls) -> ModuleType:
        import numpy.random
        return numpy.random


def get_available_numerical_packages(list_installed_bundles: tp.Optional[bool] = False) \
        -> tp.Tuple[tp.Type[NumericalPackageBundle], ...]:
    bundles = tuple(b for b in (B200Bundle,) if b.is_installed())
    if list_installed_bundles:
        print('Required packages found for the following benchmarks:')
        for b in bundles:
            print(b.label())
        print()
    return bundles
