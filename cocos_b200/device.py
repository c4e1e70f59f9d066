"""Device enumeration / selection / sync: the slice of cocos/device.py the hot path needs.

Mirrors ``ComputeDevice``, ``ComputeDeviceManager``, ``sync``, ``info``, ``init``
(cocos/device.py:39-205).  There is one backend (CUDA on B200); ArrayFire's
cpu/cuda/opencl backend switch (device.py:166-172) is out of scope.  The current
device is tracked per host thread so that ComputeDevicePool's per-GPU worker
threads each drive their own device.
"""
import ctypes
import threading
import typing as tp

from . import _lib


class ComputeDevice:
    """One GPU (cocos/device.py:39-92)."""

    def __init__(self, id: int, name: str, backend: str = "cuda", toolkit_version: str = "",
                 compute_version: str = "") -> None:
        self._id, self._name, self._backend = id, name, backend
        self._toolkit_version, self._compute_version = toolkit_version, compute_version

    id = property(lambda self: self._id)
    name = property(lambda self: self._name)
    backend = property(lambda self: self._backend)
    toolkit_version = property(lambda self: self._toolkit_version)
    compute_version = property(lambda self: self._compute_version)

    def sync(self):
        _lib.context(self._id).sync()

    def __eq__(self, other):
        return isinstance(other, ComputeDevice) and other.id == self.id

    def __hash__(self):
        return hash(("ComputeDevice", self._id))

    def __repr__(self):
        return f"ComputeDevice(id={self._id}, name={self._name!r}, compute={self._compute_version})"


class ComputeDeviceManager:
    """cocos/device.py:95-163, CUDA only."""
    _local = threading.local()
    _devices: tp.Optional[tp.Tuple[ComputeDevice, ...]] = None

    @classmethod
    def get_number_of_compute_devices(cls) -> int:
        return _lib.device_count()

    @classmethod
    def get_compute_devices(cls) -> tp.Tuple[ComputeDevice, ...]:
        if cls._devices is None:
            out = []
            for i in range(_lib.device_count()):
                buf = ctypes.create_string_buffer(256)
                _lib.check(_lib.lib().cb_device_name(i, buf, 256))
                sm, major, minor, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
                _lib.check(_lib.lib().cb_device_attributes(i, ctypes.byref(sm), ctypes.byref(major),
                                                           ctypes.byref(minor), ctypes.byref(mem)))
                out.append(ComputeDevice(i, buf.value.decode(), "cuda", "12.9", f"{major.value}.{minor.value}"))
            cls._devices = tuple(out)
        return cls._devices

    @classmethod
    def get_compute_device(cls, compute_device_id: int) -> ComputeDevice:
        devices = cls.get_compute_devices()
        if not 0 <= compute_device_id < len(devices):
            raise ValueError(f"compute device {compute_device_id} does not exist ({len(devices)} devices)")
        return devices[compute_device_id]

    @classmethod
    def get_current_compute_device_id(cls) -> int:
        return getattr(cls._local, "device", 0)

    @classmethod
    def get_current_compute_device(cls) -> ComputeDevice:
        return cls.get_compute_device(cls.get_current_compute_device_id())

    @classmethod
    def set_compute_device(cls, compute_device: tp.Union[int, ComputeDevice]):
        if isinstance(compute_device, ComputeDevice):
            compute_device = compute_device.id
        elif not isinstance(compute_device, int):
            raise TypeError("compute_device must be an int or a ComputeDevice")
        cls.get_compute_device(compute_device)          # range check
        cls._local.device = compute_device


def current_context() -> "_lib.Context":
    return _lib.context(ComputeDeviceManager.get_current_compute_device_id())


def init(backend: tp.Optional[str] = None, device_id: tp.Optional[int] = None):
    """cocos.device.init (device.py:166-172); only the CUDA backend exists here."""
    if backend not in (None, "cuda"):
        raise ValueError("cocos_b200 has a single backend: 'cuda' on NVIDIA B200")
    if device_id is not None:
        ComputeDeviceManager.set_compute_device(device_id)
    current_context()


def sync(device_id: tp.Optional[int] = None):
    """Block until the device's stream has drained (device.py:195-205)."""
    if device_id is None:
        device_id = ComputeDeviceManager.get_current_compute_device_id()
    _lib.context(device_id).sync()


def info():
    """Print the device list (device.py:175-188)."""
    print(_lib.lib().cb_version().decode())
    for d in ComputeDeviceManager.get_compute_devices():
        mark = "*" if d.id == ComputeDeviceManager.get_current_compute_device_id() else " "
        print(f"{mark}[{d.id}] {d.name} (sm_{d.compute_version.replace('.', '')})")
