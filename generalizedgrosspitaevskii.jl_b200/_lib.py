"""ctypes binding of libggp.so (include/ggp.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback on the product path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libggp.so")
LIB_PATH = os.environ.get("GGP_LIBRARY", LIB_PATH)  # A/B measurements of kernel variants (tools/)

GGP_ABI_VERSION = 5
GGP_MAX_COMPONENTS = 4
GGP_C64, GGP_C128 = 0, 1
TABLE_NONE, TABLE_SCALAR, TABLE_DIAG, TABLE_FULL, TABLE_SEP_AXES = 0, 1, 2, 3, 4
NL_NONE, NL_DIAG, NL_MATRIX = 0, 1, 2
PUMP_NONE, PUMP_SEPARABLE, PUMP_DENSE = 0, 1, 2
NOISE_NONE, NOISE_CONST, NOISE_FIELD = 0, 1, 2
OBS_DENSITY, OBS_MOMENTUM, OBS_NORM, OBS_G2_MOMENTUM = 0, 1, 2, 3

EXPORTS = [
    "ggp_version", "ggp_device_count", "ggp_last_error", "ggp_plan_create", "ggp_plan_destroy",
    "ggp_set_state", "ggp_get_state", "ggp_step", "ggp_synchronize", "ggp_observe",
    "ggp_comm_unique_id", "ggp_comm_init", "ggp_slab_ipc_export", "ggp_slab_ipc_attach", "ggp_state_device_ptr", "ggp_timer_begin", "ggp_timer_end",
    "ggp_launch_count", "ggp_host_alloc", "ggp_host_free", "ggp_device_bytes", "ggp_profile_enable",
    "ggp_profile_read", "ggp_debug_l2_flush", "ggp_debug_flush_only",
    "ggp_observe_windowed", "ggp_save_async", "ggp_save_wait", "ggp_checkpoint_bytes", "ggp_checkpoint_save", "ggp_checkpoint_load",
    "ggp_step_dense", "ggp_host_register", "ggp_host_unregister", "ggp_profile_steps_enable", "ggp_profile_steps_read",
]


class GgpDesc(C.Structure):
    """Mirror of `struct ggp_desc` (include/ggp.h) -- field order and types must match."""
    _fields_ = [
        ("abi_version", C.c_uint32), ("struct_size", C.c_uint32),
        ("ndim", C.c_int32), ("ncomp", C.c_int32),
        ("n", C.c_int64 * 3), ("nbatch", C.c_int64), ("batch_offset", C.c_int64),
        ("precision", C.c_int32), ("table_precision", C.c_int32),
        ("device", C.c_int32), ("reserved0", C.c_int32),
        ("stream", C.c_void_p),
        ("dt", C.c_double),
        ("disp_kind", C.c_int32), ("pot_kind", C.c_int32),
        ("disp_table", C.c_void_p), ("pot_table", C.c_void_p),
        ("nl_kind", C.c_int32), ("nl_scalar", C.c_int32),
        ("nl_c", (C.c_double * 2) * 2), ("nl_g", ((C.c_double * 2) * 2) * 2),
        ("pump_kind", C.c_int32), ("pump_ncomp", C.c_int32),
        ("pump_table", C.c_void_p), ("pump_amp0", C.c_double * 2),
        ("noise_kind", C.c_int32), ("noise_real", C.c_int32),
        ("noise_eta", (C.c_double * 2) * 2), ("seed", C.c_uint64),
        ("slab_nranks", C.c_int32), ("slab_rank", C.c_int32),
        ("noise_alpha", ((C.c_double * 2) * 2) * 2), ("noise_profile", C.c_void_p),
        ("disp_sep_tol", C.c_double),
        ("disp_axes", C.c_void_p * 3),
        ("mixed_precision_tables", C.c_int32), ("reserved1", C.c_int32),
        ("nl_c_ext", C.c_void_p), ("nl_g_ext", C.c_void_p),
        ("noise_eta_ext", C.c_void_p), ("noise_alpha_ext", C.c_void_p),
    ]


class GgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libggp error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libggp.so (built by __graft_entry__.build() / `make -C csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "This backend has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i64 = C.c_void_p, C.c_int64
    lib.ggp_version.restype = C.c_int
    lib.ggp_device_count.restype = C.c_int
    lib.ggp_last_error.restype = C.c_char_p
    lib.ggp_plan_create.argtypes = [C.POINTER(GgpDesc), C.POINTER(vp)]
    lib.ggp_plan_destroy.argtypes = [vp]
    lib.ggp_set_state.argtypes = [vp, C.POINTER(vp)]
    lib.ggp_get_state.argtypes = [vp, C.POINTER(vp)]
    lib.ggp_step.argtypes = [vp, i64, vp, C.POINTER(vp)]
    lib.ggp_synchronize.argtypes = [vp]
    lib.ggp_save_async.argtypes = [vp, C.POINTER(vp)]
    lib.ggp_save_wait.argtypes = [vp]
    lib.ggp_checkpoint_bytes.argtypes = [vp]
    lib.ggp_checkpoint_bytes.restype = i64
    lib.ggp_checkpoint_save.argtypes = [vp, vp, C.c_uint64]
    lib.ggp_checkpoint_load.argtypes = [vp, vp, C.c_uint64]
    lib.ggp_observe.argtypes = [vp, C.c_int, vp]
    lib.ggp_observe_windowed.argtypes = [vp, vp, vp, vp]
    lib.ggp_comm_unique_id.argtypes = [vp]
    lib.ggp_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.ggp_slab_ipc_export.argtypes = [vp, vp]
    lib.ggp_slab_ipc_attach.argtypes = [vp, vp]
    lib.ggp_state_device_ptr.argtypes = [vp, C.c_int]
    lib.ggp_state_device_ptr.restype = vp
    lib.ggp_timer_begin.argtypes = [vp]
    lib.ggp_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ggp_launch_count.argtypes = [vp]
    lib.ggp_launch_count.restype = i64
    lib.ggp_device_bytes.argtypes = [vp]
    lib.ggp_device_bytes.restype = i64
    lib.ggp_host_alloc.argtypes = [C.c_uint64]
    lib.ggp_host_alloc.restype = vp
    lib.ggp_host_free.argtypes = [vp]
    lib.ggp_profile_enable.argtypes = [vp, C.c_int]
    lib.ggp_debug_l2_flush.argtypes = [vp, C.c_uint64]
    lib.ggp_debug_flush_only.argtypes = [vp, i64, C.POINTER(C.c_float)]
    lib.ggp_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.ggp_step_dense.argtypes = [vp, i64, C.POINTER(vp), C.POINTER(vp)]
    lib.ggp_host_register.argtypes = [vp, C.c_uint64]
    lib.ggp_host_unregister.argtypes = [vp]
    lib.ggp_profile_steps_enable.argtypes = [vp, C.c_int, C.c_uint64]
    lib.ggp_profile_steps_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().ggp_last_error()
        raise GgpError(rc, msg.decode() if msg else "?")
