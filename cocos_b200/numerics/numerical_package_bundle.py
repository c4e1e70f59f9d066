"""The drop-in boundary: the backend "bundle" protocol of cocos/numerics/numerical_package_bundle.py:6-29.

The reference selects its backend by passing a ``NumericalPackageBundle`` subclass into the Heston functions
(heston_utilities.py:35,104,134) and calling five classmethods on it: ``is_installed``, ``label``, ``module``,
``random_module``, ``synchronize``.  Here a bundle is declared by three class attributes and the protocol methods
are implemented once; ``B200Bundle`` is this package's bundle and ``CocosBundle`` its alias, so call sites written
for the reference pick up the B200 path unchanged.
"""
import importlib
import os
import typing as tp
from types import ModuleType


class NumericalPackageBundle:
    """Protocol: ``is_installed() / label() / module() / random_module() / synchronize()`` (all classmethods)."""
    display_name: tp.ClassVar[str] = ""
    array_module: tp.ClassVar[str] = ""
    rng_module: tp.ClassVar[str] = ""

    @classmethod
    def is_installed(cls) -> bool:
        try:
            importlib.import_module(cls.array_module)
        except Exception:           # noqa: BLE001  (the reference swallows every import failure as well)
            return False
        return True

    @classmethod
    def label(cls) -> str:
        return cls.display_name

    @classmethod
    def module(cls) -> ModuleType:
        return importlib.import_module(cls.array_module)

    @classmethod
    def random_module(cls) -> ModuleType:
        return importlib.import_module(cls.rng_module)

    @classmethod
    def synchronize(cls) -> None:
        """Nothing to wait for on a host backend."""


class B200Bundle(NumericalPackageBundle):
    """The sm_100a kernels of this package behind the reference's bundle protocol."""
    display_name = "Cocos-B200"
    array_module = "cocos_b200.numerics"
    rng_module = "cocos_b200.numerics.random"

    @classmethod
    def is_installed(cls) -> bool:
        from .. import _lib
        return os.path.exists(_lib.LIB_PATH)          # the compiled library is the product; no fallback without it

    @classmethod
    def synchronize(cls) -> None:
        from ..device import sync
        sync()


CocosBundle = B200Bundle


class NumpyBundle(NumericalPackageBundle):
    """Kept so that code enumerating bundles still imports (reference: numerical_package_bundle.py:60-80).
    The Heston entry points of this package refuse it: there is no CPU branch here (oracle/ holds the restatement)."""
    display_name = "NumPy"
    array_module = "numpy"
    rng_module = "numpy.random"


def get_available_numerical_packages(list_installed_bundles: tp.Optional[bool] = False) \
        -> tp.Tuple[tp.Type[NumericalPackageBundle], ...]:
    """Bundles whose compute path is available here (only the B200 bundle can price)."""
    found = tuple(bundle for bundle in (B200Bundle,) if bundle.is_installed())
    if list_installed_bundles:
        print("Required packages found for the following benchmarks:")
        print("\n".join(bundle.label() for bundle in found), end="\n\n")
    return found
