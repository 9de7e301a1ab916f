"""TEST INFRASTRUCTURE: numpy-level driver of tests/simt/libsoftgrip_simt.so -- the kernel + C-ABI source of
soft-grip_b200/csrc compiled for the CPU under the SIMT emulator (tests/simt/simt.h).

Used by the ``-m "not gpu"`` tests to check the warp-synchronous kernel logic against the oracle without a GPU.
The product never imports this module and libsoftgrip.so has no host path.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE], stdout=subprocess.DEVNULL)
    return os.path.join(HERE, "libsoftgrip_simt.so")


def lib():
    global _LIB
    if _LIB is None:
        sig = importlib.import_module("soft-grip_b200._lib").SIGNATURES
        L = C.CDLL(build())
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc < 0:
        raise RuntimeError(lib().sg_last_error().decode())
    return rc


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuBatch:
    """W worlds stepping under the emulator; mirrors the parts of BatchedManEnv the parity tests use."""

    def __init__(self, blob_path, W, prec=64, lpw=8, aux_smem=0, nw=None, qv_smem=0, team=0, joint_ids=range(11, 64), tendon0=1):
        batched = importlib.import_module("soft-grip_b200.batched")
        mjcf = importlib.import_module("soft-grip_b200.mjcf")
        lib_ = importlib.import_module("soft-grip_b200._lib")
        self.L = lib()
        os.environ["SOFTGRIP_LPW"] = str(lpw)
        os.environ["SOFTGRIP_AUX_SMEM"] = str(aux_smem)
        os.environ["SOFTGRIP_QV_SMEM"] = str(qv_smem)
        os.environ["SOFTGRIP_TEAM"] = str(team)
        if nw is None:
            os.environ.pop("SOFTGRIP_NW", None)
        else:
            os.environ["SOFTGRIP_NW"] = str(nw)
        blob = open(blob_path, "rb").read()
        self.model = mjcf.load_blob(blob_path)
        self.m = C.c_void_p()
        check(self.L.sg_model_load(blob, len(blob), C.byref(self.m)))
        self.info = lib_.SgInfo()
        check(self.L.sg_model_info(self.m, C.byref(self.info)))
        mask = np.zeros(self.info.nv, dtype=np.int32)
        mask[[j for j in joint_ids if j < self.info.nv]] = 1
        check(self.L.sg_model_set_stiffness_targets(self.m, mask.ctypes.data_as(C.POINTER(C.c_int)), tendon0))
        gm = batched.geom_name_mask(self.model.names["geom"], "OBJ", ("g12", "g2"))
        check(self.L.sg_model_set_geom_mask(self.m, gm.ctypes.data_as(C.POINTER(C.c_int))))
        self.b = C.c_void_p()
        check(self.L.sg_batch_create(self.m, W, 0, prec, C.byref(self.b)))
        self.W, self.prec = W, prec
        self.dt = np.float32 if prec == 32 else np.float64
        self.nv, self.nu, self.nsd = self.info.nv, self.info.nu, self.info.nsensordata
        self.sens = np.zeros((W, self.nsd), dtype=self.dt)
        self.touch = np.zeros(W, dtype=np.int32)
        self._keep = []

    def close(self):
        if self.b:
            self.L.sg_batch_destroy(self.b); self.b = None
        if self.m:
            self.L.sg_model_destroy(self.m); self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, stiffness=None, damping=None, tdamping=None, objoff=None):
        arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (stiffness, damping, tdamping, objoff)]
        self._keep = arrs
        check(self.L.sg_batch_set_params(self.b, *[_p(a) for a in arrs], None))

    def reset(self):
        check(self.L.sg_batch_reset(self.b, None))

    def set_ctrl(self, ctrl):
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(ctrl, dtype=np.float64), (self.W, self.nu)))
        check(self.L.sg_batch_set_ctrl(self.b, _p(c), None))

    def set_state(self, q, v, act, warm):
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (self.W, n)))
                for x, n in ((q, self.nv), (v, self.nv), (act, self.nu), (warm, self.nv))]
        D = C.POINTER(C.c_double)
        check(self.L.sg_batch_set_state(self.b, *[a.ctypes.data_as(D) for a in arrs]))

    def get_state(self):
        q = np.zeros((self.W, self.nv)); v = np.zeros((self.W, self.nv)); a = np.zeros((self.W, self.nu)); w = np.zeros((self.W, self.nv))
        D = C.POINTER(C.c_double)
        check(self.L.sg_batch_get_state(self.b, *[x.ctypes.data_as(D) for x in (q, v, a, w)]))
        return q, v, a, w

    def step(self, nsub=1):
        check(self.L.sg_batch_step(self.b, int(nsub), _p(self.sens), _p(self.touch), None))
        return self.sens.astype(np.float64), self.touch.copy()

    def forward(self):
        check(self.L.sg_batch_forward(self.b, _p(self.sens), _p(self.touch), None))
        return self.sens.astype(np.float64), self.touch.copy()

    def rollout(self, schedule, sim_start=1, sim_step=7, want_touch=True):
        lib_ = importlib.import_module("soft-grip_b200._lib")
        ev, val = schedule
        ev = np.ascontiguousarray(ev, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
        T = ev.shape[0]
        sc = lib_.SgSchedule(sim_start, sim_step, T, ev.ctypes.data_as(C.POINTER(C.c_int)), val.ctypes.data_as(C.POINTER(C.c_double)))
        traj = np.zeros((self.W, T, self.nsd), dtype=self.dt)
        touch = np.zeros((self.W, T), dtype=np.int32) if want_touch else None
        check(self.L.sg_batch_rollout(self.b, C.byref(sc), _p(traj), _p(touch), None))
        return traj.astype(np.float64), touch, self.status()

    def status(self, clear=False):
        out = np.zeros(self.W, dtype=np.int32)
        check(self.L.sg_batch_status(self.b, out.ctypes.data_as(C.POINTER(C.c_int)), int(clear)))
        return out

    def set_debug_world(self, w):
        check(self.L.sg_batch_set_debug_world(self.b, int(w)))

    def debug(self, key):
        n = check(self.L.sg_batch_debug_get(self.b, key.encode(), None, 0))
        out = np.zeros(max(n, 1))
        check(self.L.sg_batch_debug_get(self.b, key.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return out[:n]
