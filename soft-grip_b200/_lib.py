"""ctypes loader of libsoftgrip.so (C-ABI in include/softgrip.h).

There is no CPU path: if the library is missing or no CUDA device is present the calls fail loudly.
Build with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C soft-grip_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SOFTGRIP_LIB: development knob for A/B runs of kernel variants (csrc/Makefile `variant`); never set by tests/ or bench.py
LIB_PATH = os.environ.get("SOFTGRIP_LIB") or os.path.join(_HERE, "libsoftgrip.so")
_LIB = None


class SoftGripError(RuntimeError):
    pass


class SgInfo(C.Structure):
    _fields_ = [("nv", C.c_int), ("nfinger", C.c_int), ("nshell", C.c_int), ("neq", C.c_int), ("nu", C.c_int),
                ("nsensordata", C.c_int), ("nlevels", C.c_int), ("maxcon", C.c_int), ("smem_bytes32", C.c_int),
                ("smem_bytes64", C.c_int), ("ngeom", C.c_int), ("reserved", C.c_int * 5)]


class SgSchedule(C.Structure):
    _fields_ = [("sim_start", C.c_int), ("sim_step", C.c_int), ("nrows", C.c_int),
                ("ctrl_event", C.POINTER(C.c_int)), ("ctrl_value", C.POINTER(C.c_double))]


# every symbol include/softgrip.h declares, with its signature (tests check the .so exports all of them)
P, D, I, V = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
SIGNATURES = {
    "sg_last_error": (C.c_char_p, []),
    "sg_version": (C.c_int, []),
    "sg_model_load": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(P)]),
    "sg_model_info": (C.c_int, [P, C.POINTER(SgInfo)]),
    "sg_model_destroy": (None, [P]),
    "sg_batch_create": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.POINTER(P)]),
    "sg_batch_destroy": (None, [P]),
    "sg_batch_nworlds": (C.c_int, [P]),
    "sg_batch_precision": (C.c_int, [P]),
    "sg_model_set_stiffness_targets": (C.c_int, [P, I, C.c_int]),
    "sg_batch_set_params": (C.c_int, [P, V, V, V, V, V]),
    "sg_model_set_geom_mask": (C.c_int, [P, I]),
    "sg_batch_reset": (C.c_int, [P, V]),
    "sg_batch_set_ctrl": (C.c_int, [P, V, V]),
    "sg_batch_set_ctrl_all": (C.c_int, [P, D, V]),
    "sg_batch_forward": (C.c_int, [P, V, V, V]),
    "sg_batch_step": (C.c_int, [P, C.c_int, V, V, V]),
    "sg_batch_rollout": (C.c_int, [P, C.POINTER(SgSchedule), V, V, V]),
    "sg_batch_rollout_host": (C.c_int, [P, C.POINTER(SgSchedule), V, V, V, V]),
    "sg_batch_rollout_host_params": (C.c_int, [P, C.POINTER(SgSchedule), V, V, V, V, V, V, V]),
    "sg_batch_set_traj_layout": (C.c_int, [P, C.c_int]),
    "sg_batch_get_state": (C.c_int, [P, D, D, D, D]),
    "sg_batch_set_state": (C.c_int, [P, D, D, D, D]),
    "sg_batch_status": (C.c_int, [P, I, C.c_int]),
    "sg_batch_sync": (C.c_int, [P, V]),
    "sg_batch_set_debug_world": (C.c_int, [P, C.c_int]),
    "sg_batch_debug_get": (C.c_int, [P, C.c_char_p, D, C.c_int]),
    "sg_batch_launch_count": (C.c_longlong, [P]),
    "sg_batch_config": (C.c_int, [P, I]),
    "sg_batch_prof_get": (C.c_int, [P, C.POINTER(C.c_ulonglong), C.c_int]),
    "sg_traj_add_noise": (C.c_int, [V, V, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_double, C.c_double, C.c_ulonglong, V, V,
                                    C.c_int, C.c_int, V]),
    "sg_traj_channel_stats": (C.c_int, [V, C.c_longlong, C.c_int, C.c_int, C.c_int, V, V, V, C.c_longlong, V]),
    "sg_traj_stats_workspace_bytes": (C.c_longlong, [C.c_longlong, C.c_int, C.c_int]),
    "sg_traj_mask_contact": (C.c_int, [V, V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, V, C.c_int, C.c_int, V]),
}


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SoftGripError("libsoftgrip.so is not built (%s); run __graft_entry__.build() -- there is no CPU fallback"
                                % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc < 0:
        raise SoftGripError(lib().sg_last_error().decode() or "softgrip error %d" % rc)
    return rc
