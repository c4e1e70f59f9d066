"""``cocos_b200.multi_processing``: map-reduce over GPUs (streams + NCCL) and CPU cores."""
from .device_pool import ComputeDevicePool                                              # noqa: F401
from .multi_core_batch_processing import map_combine_multicore, map_reduce_multicore    # noqa: F401
from .single_device_batch_processing import map_combine_single_device, map_reduce_single_device  # noqa: F401
from .utilities import (MultiprocessingPoolType, generate_slices_with_batch_size,      # noqa: F401
                        generate_slices_with_number_of_batches)
