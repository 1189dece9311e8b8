"""ctypes binding of the C ABI in include/mppi_b200.h (libmppi_b200.so).

The shared library is the product; this module only loads it and declares the
prototypes.  There is no Python or CPU implementation behind it: if the
library is missing, importing fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MPPI_B200_LIB: development aid (A/B builds of the library under tools/_build)
LIB_PATH = os.environ.get("MPPI_B200_LIB") or os.path.join(HERE, "libmppi_b200.so")

MAX_ACT = 4
COMM_ID_BYTES = 128

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_COMM, ERR_STATE = -1, -2, -3, -4, -5

FLAG_STRICT_ARITH = 1 << 0
FLAG_INJECTED_NOISE = 1 << 1
FLAG_CLAMP_ACTIONS = 1 << 2
FLAG_REINIT_INIT_ACT = 1 << 3
FLAG_NO_GRAPH = 1 << 4
FLAG_FUSED_SAMPLING = 1 << 5
FLAG_SPLIT_KERNELS = 1 << 6
FLAG_STEP_KERNEL = 1 << 7
FLAG_AUTO_CHAIN = 1 << 8
FLAG_PIPELINED_SAMPLING = 1 << 9
FLAG_TILE_KERNEL = 1 << 10

COMM_NONE, COMM_NCCL, COMM_P2P = 0, 1, 2
MODEL_POINT_MASS, MODEL_LINEAR_AXIS = 0, 1
P2P_HANDLE_BYTES = 64

K_SAMPLE, K_ROLLOUT, K_COMM_MIN, K_WEIGHTS, K_AVERAGE, K_COMM_SUM, K_FINALIZE, K_COUNT = range(8)

# every symbol include/mppi_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "mppi_params_default", "mppi_create", "mppi_create_multi", "mppi_destroy", "mppi_set_problem",
    "mppi_set_terminal_weights", "mppi_set_state",
    "mppi_step", "mppi_step_enqueue", "mppi_step_wait", "mppi_get_u", "mppi_set_u",
    "mppi_get_info", "mppi_get_flags", "mppi_get_step_info", "mppi_set_noise", "mppi_set_noise_mode",
    "mppi_sample_only", "mppi_shard_range", "mppi_local_samples", "mppi_chain_estimate", "mppi_timer_start", "mppi_timer_stop",
    "mppi_set_profiling", "mppi_get_kernel_times", "mppi_get_exchange_times", "mppi_get_launch_count", "mppi_kernel_name",
    "mppi_comm_unique_id", "mppi_comm_p2p_handle", "mppi_comm_p2p_connect", "mppi_last_error",
    "mppi_abi_version",
]


class MppiParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("samples", C.c_int64),
        ("horizon", C.c_int32),
        ("state_dim", C.c_int32),
        ("act_dim", C.c_int32),
        ("verbose", C.c_int32),
        ("dt", C.c_float),
        ("lambda_", C.c_float),
        ("sigma", C.c_float * MAX_ACT),
        ("inv_sigma", C.c_float * MAX_ACT),
        ("init_act", C.c_float * MAX_ACT),
        ("max_act", C.c_float * MAX_ACT),
        ("seed", C.c_uint64),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("comm", C.c_int32),
        ("comm_id", C.c_uint8 * COMM_ID_BYTES),
        ("model", C.c_int32),
        ("state_gain", C.c_float * 4),
        ("act_gain", C.c_float * 2),
        ("philox_rounds", C.c_int32),
    ]


class MppiStepInfo(C.Structure):
    _fields_ = [("beta", C.c_float), ("eta", C.c_float), ("argmin", C.c_int64),
                ("step", C.c_uint64)]


class MppiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mppi_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libmppi_b200.so (no GPU needed to load; mppi_create needs one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C mppi_gpu_b200/csrc`. There is no fallback implementation.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    fp = C.c_void_p   # float* passed as raw addresses (numpy .ctypes.data) or None
    L.mppi_abi_version.restype = C.c_int
    L.mppi_last_error.restype = C.c_char_p
    L.mppi_kernel_name.restype = C.c_char_p
    L.mppi_kernel_name.argtypes = [C.c_int]
    L.mppi_params_default.argtypes = [C.POINTER(MppiParams)]
    L.mppi_create.argtypes = [C.POINTER(MppiParams), C.POINTER(H)]
    L.mppi_create_multi.argtypes = [C.POINTER(MppiParams), C.POINTER(C.c_int), C.c_int, C.POINTER(H)]
    L.mppi_destroy.argtypes = [H]
    L.mppi_set_problem.argtypes = [H, fp, fp, fp, fp]
    L.mppi_set_terminal_weights.argtypes = [H, fp]
    L.mppi_chain_estimate.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                      C.POINTER(C.c_uint32)]
    L.mppi_set_state.argtypes = [H, fp]
    L.mppi_step.argtypes = [H, fp]
    L.mppi_step_enqueue.argtypes = [H]
    L.mppi_step_wait.argtypes = [H, fp]
    L.mppi_get_u.argtypes = [H, fp]
    L.mppi_set_u.argtypes = [H, fp]
    L.mppi_get_info.argtypes = [H, fp, fp, fp, fp, fp, fp, fp]
    L.mppi_get_flags.argtypes = [H, C.POINTER(C.c_uint32)]
    L.mppi_get_step_info.argtypes = [H, C.POINTER(MppiStepInfo)]
    L.mppi_set_noise.argtypes = [H, fp]
    L.mppi_set_noise_mode.argtypes = [H, C.c_int]
    L.mppi_sample_only.argtypes = [H, C.c_uint64]
    L.mppi_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64)]
    L.mppi_local_samples.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mppi_timer_start.argtypes = [H]
    L.mppi_timer_stop.argtypes = [H, C.POINTER(C.c_float)]
    L.mppi_set_profiling.argtypes = [H, C.c_int]
    L.mppi_get_kernel_times.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.mppi_get_launch_count.argtypes = [H, C.POINTER(C.c_int64)]
    L.mppi_get_exchange_times.argtypes = [H, C.POINTER(C.c_double)]
    L.mppi_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.mppi_comm_p2p_handle.argtypes = [H, C.POINTER(C.c_uint8)]
    L.mppi_comm_p2p_connect.argtypes = [H, C.POINTER(C.c_uint8)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("mppi_last_error", "mppi_kernel_name"):
            fn.restype = C.c_int
    if L.mppi_abi_version() != 3:
        raise ImportError("libmppi_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise MppiError(rc, load().mppi_last_error().decode(errors="replace"))


def shard_range(samples, rank, world_size):
    b, e = C.c_int64(), C.c_int64()
    check(load().mppi_shard_range(int(samples), int(rank), int(world_size), C.byref(b), C.byref(e)))
    return int(b.value), int(e.value)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def chain_estimate(samples_local, horizon, act_dim, num_sms=148):
    """(est_us = {unfused, fused, step}, chosen flag) of MPPI_FLAG_AUTO_CHAIN for a shard; host only."""
    est = (C.c_double * 3)()
    ch = C.c_uint32(0)
    check(load().mppi_chain_estimate(int(samples_local), int(horizon), int(act_dim), int(num_sms), est,
                                     C.byref(ch)))
    return {"unfused": est[0], "fused": est[1], "step": est[2]}, int(ch.value)
