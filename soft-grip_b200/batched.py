"""BatchedManEnv: W independent copies of the reference's ManEnv on one B200.

The batched variant of ``environment.ManEnv`` (ref: environment/manenv.py:8-126): same verbs
(``reset / step / close_hand / loose_hand / toggle_grip / set_new_stiffness``) over a batch of worlds,
returning torch tensors, plus ``rollout`` which runs the whole squeeze episode of
``create_dataset.log_into_file`` (ref: create_dataset.py:33-60) in one kernel launch.

All physics runs in libsoftgrip.so (hand-written sm_100a CUDA behind include/softgrip.h).  PyTorch is
only used for device buffers and the current stream.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import mjcf
from ._lib import SgInfo, SgSchedule, SoftGripError, check, lib

ST_DIVERGED, ST_CON_FULL, ST_EFC_FULL, ST_UNSUPPORTED = 1, 2, 4, 8
TOUCH_ANY = 1 << 30

# episode constants of the reference driver (ref: create_dataset.py:14-17)
NUM_EPISODES, MAX_ITER_PER_EP, OPEN_CLOSE_DIV, START_STEP = 1, 160, 80, 40


def world_uniform(seed, world_ids, lo, hi, stream=0):
    """Counter-based U(lo,hi) per *global* world id (splitmix64 of (seed, stream, id)): the draw of world w
    does not depend on how worlds are sharded over GPUs."""
    with np.errstate(over="ignore"):
        x = (np.asarray(world_ids, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x = x ^ (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(stream) * np.uint64(0x94D049BB133111EB))
        for _ in range(2):
            x = x ^ (x >> np.uint64(30))
            x = x * np.uint64(0xBF58476D1CE4E5B9)
            x = x ^ (x >> np.uint64(27))
            x = x * np.uint64(0x94D049BB133111EB)
            x = x ^ (x >> np.uint64(31))
    u = (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def geom_name_mask(geom_names, obj_name, finger_names):
    """bit0: name contains obj_name; bit(1+k): contains finger_names[k] (ref: manenv.py:73-78)."""
    out = np.zeros(len(geom_names), dtype=np.int32)
    for i, n in enumerate(geom_names):
        if n is None:
            continue
        if obj_name in n:
            out[i] |= 1
        for k, f in enumerate(finger_names):
            if f in n:
                out[i] |= 2 << k
    return out


def load_model(path_or_model):
    """MJCF path, compiled ``.sgm`` blob path, or an already compiled :class:`mjcf.Model`."""
    if isinstance(path_or_model, mjcf.Model):
        return path_or_model
    p = os.fspath(path_or_model)
    if p.endswith(".xml"):
        return mjcf.load_mjcf(p)
    return mjcf.load_blob(p)


def default_schedule(nu=2, n_settle=START_STEP, n_iter=MAX_ITER_PER_EP, open_close_div=OPEN_CLOSE_DIV, ctrl_mag=0.2):
    """ctrl events of create_dataset.log_into_file: ctrl=0 for the settle steps, close_hand() at row n_settle,
    toggle_grip() whenever i % open_close_div == 0 and i > 0 (ref: create_dataset.py:41-60)."""
    T = n_settle + n_iter
    ev = np.zeros(T, dtype=np.int32)
    val = np.zeros((T, nu), dtype=np.float64)
    closing = True
    if n_iter > 0:
        ev[n_settle] = 1
        val[n_settle] = -ctrl_mag
    for i in range(n_iter):
        if open_close_div > 0 and i % open_close_div == 0 and i > 0:
            closing = not closing
            ev[n_settle + i] = 1
            val[n_settle + i] = -ctrl_mag if closing else ctrl_mag
    return ev, val


class DeviceModel:
    """sg_model handle + the compiled tables it came from."""

    def __init__(self, model, joint_ids=range(11, 64), tendon_ids=(0,), obj_name="OBJ", finger_names=("g12", "g2")):
        self.model = load_model(model)
        self.L = lib()
        self.blob = mjcf.model_to_blob(self.model)
        h = C.c_void_p()
        check(self.L.sg_model_load(self.blob, len(self.blob), C.byref(h)))
        self.h = h
        info = SgInfo()
        check(self.L.sg_model_info(self.h, C.byref(info)))
        self.info = info
        mask = np.zeros(info.nv, dtype=np.int32)
        ids = [j for j in joint_ids if j < info.nv]
        mask[ids] = 1
        self.joint_mask = mask
        self.tendon0 = int(0 in tuple(tendon_ids))
        check(self.L.sg_model_set_stiffness_targets(self.h, mask.ctypes.data_as(C.POINTER(C.c_int)), self.tendon0))
        self.finger_names = tuple(finger_names)
        gm = geom_name_mask(self.model.names["geom"], obj_name, self.finger_names)
        check(self.L.sg_model_set_geom_mask(self.h, gm.ctypes.data_as(C.POINTER(C.c_int))))
        self.all_fingers = (1 << len(self.finger_names)) - 1

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sg_model_destroy(self.h)
            self.h = None


class BatchedManEnv:
    """W worlds of one gripper+object model stepping in lock-step on one GPU."""

    joint_ids = list(range(11, 64))      # ref: manenv.py:12
    tendon_ids = list(range(1))          # ref: manenv.py:13
    finger_names = ['g12', 'g2']         # ref: manenv.py:17
    obj_name = 'OBJ'                     # ref: manenv.py:18

    def __init__(self, env_path, num_worlds, device="cuda:0", dtype=None, seed=0, sim_start=1, sim_step=7,
                 world_offset=0, contact_mode="intended"):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise SoftGripError("BatchedManEnv needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise SoftGripError("BatchedManEnv only runs on CUDA devices")
        dtype = torch.float32 if dtype is None else dtype
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 (fast path) or torch.float64 (verification build)")
        self.dtype = dtype
        self.W = int(num_worlds)
        self.seed, self.world_offset = int(seed), int(world_offset)
        self.sim_start, self.sim_step = int(sim_start), int(sim_step)
        if contact_mode not in ("intended", "reference"):
            raise ValueError("contact_mode must be 'intended' or 'reference'")
        self.contact_mode = contact_mode
        self.dm = env_path if isinstance(env_path, DeviceModel) else DeviceModel(
            env_path, self.joint_ids, self.tendon_ids, self.obj_name, self.finger_names)
        self.L = self.dm.L
        self.info = self.dm.info
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._device_index = int(idx)
        check(self.L.sg_batch_create(self.dm.h, self.W, idx, 32 if dtype == torch.float32 else 64, C.byref(h)))
        self.h = h
        self.nsd, self.nu, self.nv = self.info.nsensordata, self.info.nu, self.info.nv
        self.is_closing = True
        self.episode = 0
        self.stiffness = torch.full((self.W,), 700.0, dtype=torch.float64, device=self.device)
        self._fingers_left = torch.full((self.W,), self.dm.all_fingers, dtype=torch.int32, device=self.device)
        self._sens = torch.zeros((self.W, self.nsd), dtype=dtype, device=self.device)
        self._touch = torch.zeros((self.W,), dtype=torch.int32, device=self.device)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sg_batch_destroy(self.h)
            self.h = None

    # ---- plumbing -------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _f64(self, x, shape):
        t = self.torch.as_tensor(x, dtype=self.torch.float64, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
        return t

    # ---- parameters -----------------------------------------------------------------------
    def set_new_stiffness(self, range_min=300, range_max=1400, stiffness=None):
        """Per-world stiffness on ``joint_ids`` + ``tendon_ids`` (ref: manenv.py:103-109).  Draws are
        counter-based per global world id and episode, so they do not depend on the GPU sharding."""
        if stiffness is None:
            ids = np.arange(self.world_offset, self.world_offset + self.W, dtype=np.uint64)
            k = world_uniform(self.seed, ids, float(range_min), float(range_max), stream=self.episode)
            stiffness = self.torch.from_numpy(k)
        self.stiffness = self._f64(stiffness, (self.W,))
        self._push_params()
        return self.stiffness

    def set_params(self, damping=None, tendon_damping=None, object_offset=None):
        """Extensions that are not in the reference (BASELINE.json configs[2]): per-world shell damping,
        volume-tendon damping and a translation of the object body."""
        self._damping = None if damping is None else self._f64(damping, (self.W,))
        self._tdamping = None if tendon_damping is None else self._f64(tendon_damping, (self.W,))
        self._objoff = None if object_offset is None else self._f64(object_offset, (self.W, 3))
        self._push_params()

    _damping = _tdamping = _objoff = None

    def _push_params(self):
        check(self.L.sg_batch_set_params(self.h, self._ptr(self.stiffness), self._ptr(self._damping), self._ptr(self._tdamping),
                                         self._ptr(self._objoff), self._stream()))

    # ---- ManEnv verbs ---------------------------------------------------------------------
    def reset(self, stiffness=None):
        """ManEnv.reset (ref: manenv.py:55-63): new stiffness, mj_resetData, mj_forward, ``sim_start`` steps."""
        k = self.set_new_stiffness(stiffness=stiffness)
        self.episode += 1
        check(self.L.sg_batch_reset(self.h, self._stream()))
        check(self.L.sg_batch_forward(self.h, self._ptr(self._sens), self._ptr(self._touch), self._stream()))
        if self.sim_start > 0:
            self.step(self.sim_start)
        return k

    def set_ctrl(self, ctrl):
        c = self._f64(ctrl, (self.W, self.nu))
        check(self.L.sg_batch_set_ctrl(self.h, self._ptr(c), self._stream()))

    def _ctrl_all(self, value):
        v = (C.c_double * self.nu)(*([float(value)] * self.nu))
        check(self.L.sg_batch_set_ctrl_all(self.h, v, self._stream()))

    def close_hand(self):
        self._ctrl_all(-0.2)             # ref: manenv.py:93-96
        self.is_closing = True

    def loose_hand(self):
        self._ctrl_all(0.2)              # ref: manenv.py:98-101
        self.is_closing = False

    def toggle_grip(self):
        if self.is_closing:
            self.loose_hand()
        else:
            self.close_hand()

    def step(self, num_steps=-1):
        """``num_steps`` physics steps (default ``sim_step``) then the sensor / contact read-out
        (ref: manenv.py:44-53,65-85).  Returns (readings[W,12], contact[W] bool)."""
        if num_steps < 1:
            num_steps = self.sim_step
        check(self.L.sg_batch_step(self.h, int(num_steps), self._ptr(self._sens), self._ptr(self._touch), self._stream()))
        return self._sens.clone(), self._contact_flag(self._touch)

    def get_sensor_sensordata(self):
        return self._sens.clone(), self._contact_flag(self._touch)

    def _contact_flag(self, touch):
        allf = self.dm.all_fingers
        if self.contact_mode == "intended":
            return (touch & allf) == allf
        # reference-literal: the class-level finger list is consumed once and never refilled, after which
        # the flag degenerates to ncon >= 1 (ref: manenv.py:70,77-83)
        was_empty = self._fingers_left == 0
        self._fingers_left = self._fingers_left & ~(touch & allf)
        return self.torch.where(was_empty, (touch & TOUCH_ANY) != 0, self._fingers_left == 0)

    def status(self, clear=False):
        out = np.zeros(self.W, dtype=np.int32)
        check(self.L.sg_batch_status(self.h, out.ctypes.data_as(C.POINTER(C.c_int)), int(clear)))
        return out

    def rollout(self, schedule=None, stiffness=None, return_touch=False, layout="WTC"):
        """Whole episode(s) on-chip: reset + ``sim_start`` steps, then one recorded row per env-step.
        Returns (traj, stiffness[W], status[W] int32[, touch[W,T]]).  ``layout="WTC"``: traj[W,T,12], one world's (T,12)
        sample contiguous (the reference's sample shape, ref: create_dataset.py:62-63); ``layout="TCW"``: traj[T,12,W],
        structure of arrays with the world index fastest."""
        torch = self.torch
        if layout not in ("WTC", "TCW"):
            raise ValueError("layout must be 'WTC' or 'TCW'")
        k = self.set_new_stiffness(stiffness=stiffness)
        self.episode += 1
        ev, val = schedule if schedule is not None else default_schedule(self.nu)
        ev = np.ascontiguousarray(ev, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        T = ev.shape[0]
        sc = SgSchedule(self.sim_start, self.sim_step, T, ev.ctypes.data_as(C.POINTER(C.c_int)),
                        val.ctypes.data_as(C.POINTER(C.c_double)))
        shape = (self.W, T, self.nsd) if layout == "WTC" else (T, self.nsd, self.W)
        traj = torch.empty(shape, dtype=self.dtype, device=self.device)
        touch = torch.empty((self.W, T), dtype=torch.int32, device=self.device) if return_touch else None
        self.status(clear=True)
        check(self.L.sg_batch_set_traj_layout(self.h, 0 if layout == "WTC" else 1))
        try:
            check(self.L.sg_batch_rollout(self.h, C.byref(sc), self._ptr(traj), self._ptr(touch), self._stream()))
        finally:
            check(self.L.sg_batch_set_traj_layout(self.h, 0))
        st = torch.from_numpy(self.status()).to(self.device)
        return (traj, k, st, touch) if return_touch else (traj, k, st)

    def mask_contact(self, traj, touch):
        """``--mask-contact`` of the reference driver for a whole rollout, on the device and in place: rows recorded
        without finger-object contact are zeroed (ref: create_dataset.py:43-44,57-58), with this environment's
        ``contact_mode`` ("reference": the aliased finger list of manenv.py:70,80 is carried from row to row and from
        episode to episode in ``_fingers_left``).  traj [W,T,C] and touch [W,T] as returned by :meth:`rollout`."""
        if traj.shape[:2] != touch.shape or traj.shape[0] != self.W or not traj.is_contiguous() or not touch.is_contiguous():
            raise ValueError("traj must be [W,T,C] and touch [W,T], both contiguous")
        if touch.dtype != self.torch.int32 or traj.dtype != self.dtype:
            raise TypeError("touch must be int32 and traj in the batch precision")
        mode = 1 if self.contact_mode == "reference" else 0
        check(self.L.sg_traj_mask_contact(self._ptr(traj), self._ptr(touch), self.W, int(traj.shape[1]), int(traj.shape[2]),
                                          self.dm.all_fingers, TOUCH_ANY, mode, self._ptr(self._fingers_left) if mode else None,
                                          32 if self.dtype == self.torch.float32 else 64,
                                          self._device_index, self._stream()))
        return traj

    # ---- state access (parity tests) ---------------------------------------------------------
    def get_state(self):
        q = np.zeros((self.W, self.nv)); v = np.zeros((self.W, self.nv)); a = np.zeros((self.W, self.nu)); w = np.zeros((self.W, self.nv))
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        check(self.L.sg_batch_get_state(self.h, dp(q), dp(v), dp(a), dp(w)))
        return q, v, a, w

    def set_state(self, qpos=None, qvel=None, act=None, warm=None):
        def prep(x, n):
            if x is None:
                return None, None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (self.W, n)))
            return a, a.ctypes.data_as(C.POINTER(C.c_double))
        keep = [prep(qpos, self.nv), prep(qvel, self.nv), prep(act, self.nu), prep(warm, self.nv)]
        check(self.L.sg_batch_set_state(self.h, *[k[1] for k in keep]))

    def debug(self, world, key):
        """Diagnostics of the last physics step of one world (set the debug world *before* stepping)."""
        n = check(self.L.sg_batch_debug_get(self.h, key.encode(), None, 0))
        out = np.zeros(max(n, 1))
        check(self.L.sg_batch_debug_get(self.h, key.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return out[:n]

    def set_debug_world(self, world):
        check(self.L.sg_batch_set_debug_world(self.h, int(world)))

    def config(self):
        """How the worlds were packed onto the SMs (sg_batch_config)."""
        out = (C.c_int * 8)()
        check(self.L.sg_batch_config(self.h, out))
        keys = ("lanes_per_world", "warps_per_cta", "worlds_per_cta", "ctas_per_sm", "smem_per_cta", "smem_per_world", "team_mode", "kernel")
        cfg = dict(zip(keys, [int(x) for x in out]))
        try:
            # where the equality rows and the contact records live during the solve
            tm = self.debug(0, "tensor_memory")
            cfg["tensor_memory_columns"] = int(tm[0])
            cfg["record_ring"] = int(tm[2])
        except Exception:  # noqa: BLE001 -- diagnostics only
            pass
        return cfg

    PHASES = ("gripper", "collide", "rows", "warmstart", "pgs_setup", "pgs_equality", "pgs_chain", "sensors", "euler", "other")

    def phase_cycles(self):
        """Development aid (SOFTGRIP_PROF=1 at creation): SM cycles per phase of the step kernel, summed over warps."""
        out = (C.c_ulonglong * 16)()
        check(self.L.sg_batch_prof_get(self.h, out, 16))
        return dict(zip(self.PHASES, [int(x) for x in out]))

    def launch_count(self):
        return int(self.L.sg_batch_launch_count(self.h))

    def synchronize(self):
        check(self.L.sg_batch_sync(self.h, self._stream()))
