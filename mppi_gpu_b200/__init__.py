"""mppi_gpu_b200 -- B200-native MPPI control-step core (sm_100a CUDA behind a C ABI).

Package contents: csrc/ (CUDA kernels + C ABI -> libmppi_b200.so), capi.py (ctypes
prototypes) and controller.py (host-side mirror of the reference's PointMassModel).
"""
from .capi import MppiError, load  # noqa: F401
from .controller import PointMassModel, comm_unique_id  # noqa: F401
