"""Drop-in replacement of the reference's ``environment.ManEnv`` (ref: environment/manenv.py:8-126).

Same constructor, class attributes, methods, return values and quirks; the physics behind it is a
one-world batch of the B200 kernels (libsoftgrip.so) instead of ``mujoco_py.MjSim``.  The fp64
verification build is used so that a single environment reproduces the CPU oracle to ~1e-12.

Differences that are deliberate and documented:
  * ``step()`` returns an explicit ``dtype=object`` array of length 2 (the reference relies on NumPy < 1.24
    building one implicitly, ref: manenv.py:53);
  * ``render()`` is a no-op, there is no viewer;
  * ``get_env()`` returns a small handle exposing the ``MjSim`` attributes ManEnv itself touches.
The contact-flag aliasing quirk (ref: manenv.py:70,77-83) is reproduced literally, including the fact that
it leaks across episodes and instances through the class attribute.
"""
import importlib
import os
import sys

import numpy as np

from .interface import Env

_pkg_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pkg_name = os.path.basename(_pkg_dir)
if os.path.dirname(_pkg_dir) not in sys.path:
    sys.path.insert(0, os.path.dirname(_pkg_dir))
_batched = importlib.import_module(_pkg_name + ".batched")

DEFAULT_DAMPING = 200


class _Data(object):
    def __init__(self, nu, nsd):
        self.ctrl = np.zeros(nu)
        self.sensordata = np.zeros(nsd)
        self.ncon = 0


class _Model(object):
    def __init__(self, model, geom_names):
        self.jnt_stiffness = model.arrays["jnt_stiffness"].copy()
        self.tendon_stiffness = model.arrays["tendon_stiffness"].copy()
        self._geom_names = geom_names

    def geom_id2name(self, gid):
        return self._geom_names[gid]


class _Sim(object):
    """What ``ManEnv.get_env()`` hands out: ``step/reset/forward`` and ``data.ctrl``, ``data.sensordata``,
    ``data.ncon``, ``model.jnt_stiffness``, ``model.tendon_stiffness`` (the MjSim subset of SURVEY section 1)."""

    def __init__(self, env_path, joint_ids, tendon_ids, obj_name, finger_names):
        import torch
        self._torch = torch
        self._joint_ids, self._tendon_ids = list(joint_ids), list(tendon_ids)
        self._dm = _batched.DeviceModel(env_path, joint_ids, tendon_ids, obj_name, tuple(finger_names))
        self._b = _batched.BatchedManEnv(self._dm, 1, dtype=torch.float64, sim_start=0, sim_step=1)
        self.model = _Model(self._dm.model, self._dm.model.names["geom"])
        self.data = _Data(self._b.nu, self._b.nsd)
        self._base_jnt = self.model.jnt_stiffness.copy()
        self._base_ten = self.model.tendon_stiffness.copy()
        self.touch = 0
        self.diverged = False

    def _push(self):
        # the only model edits ManEnv performs: one value on joint_ids and tendon_ids (ref: manenv.py:103-109)
        js, ts = self.model.jnt_stiffness, self.model.tendon_stiffness
        ids = [j for j in self._joint_ids if j < js.shape[0]]
        k = js[ids[0]] if ids else ts[self._tendon_ids[0]]
        other = np.ones(js.shape[0], dtype=bool)
        other[ids] = False
        if not (np.all(js[ids] == k) and np.array_equal(js[other], self._base_jnt[other])
                and all(ts[t] == k for t in self._tendon_ids)):
            raise NotImplementedError("only ManEnv.set_new_stiffness-style edits of jnt_stiffness/tendon_stiffness are supported")
        self._b.set_new_stiffness(stiffness=[float(k)])
        self._b.set_ctrl(self.data.ctrl.reshape(1, -1))

    def _pull(self, sens, touch):
        self.data.sensordata = sens[0].double().cpu().numpy().copy()
        self.touch = int(touch[0].item())
        self.data.ncon = 1 if (self.touch & _batched.TOUCH_ANY) else 0

    def step(self):
        self._push()
        self._b.step(1)
        self._pull(self._b._sens, self._b._touch)
        if int(self._b.status(clear=True)[0]) & _batched.ST_DIVERGED:
            self.data.ctrl[:] = 0            # mj_resetData
            self.diverged = True

    def forward(self):
        self._push()
        b = self._b
        _batched.check(b.L.sg_batch_forward(b.h, b._ptr(b._sens), b._ptr(b._touch), b._stream()))
        self._pull(b._sens, b._touch)

    def reset(self):
        _batched.check(self._b.L.sg_batch_reset(self._b.h, self._b._stream()))
        self.data.ctrl[:] = 0
        self.data.sensordata[:] = 0
        self.data.ncon = 0
        self.touch = 0


class MujocoException(Exception):
    """Stand-in for mujoco_py.builder.MujocoException (ref: manenv.py:50)."""


class ManEnv(Env):
    # ADJUST VARIABLES DEPENDING ON YOUR DATASET (ref: manenv.py:9-18)
    joint_ids = list(range(11, 64))  # JOINT INDEXES FOR 2 FINGER GRIPPER
    tendon_ids = list(range(1))
    finger_names = ['g12', 'g2']  # FINGER NAMES FOR 2 FINGER GRIPPER
    obj_name = 'OBJ'
    # pristine copy: `finger_names` itself is consumed by get_sensor_sensordata (reference quirk, see below)
    _finger_names0 = ('g12', 'g2')

    def __init__(self, sim_start, sim_step, env_paths, is_vis=True):
        super().__init__(sim_start, sim_step)
        assert len(env_paths) > 0
        self.is_vis = is_vis
        self.env_paths = env_paths
        self.env = self._make(env_paths[0])
        self.is_closing = True

    def _make(self, path):
        return _Sim(path, self.joint_ids, self.tendon_ids, self.obj_name, self._finger_names0)

    def load_env(self, num):
        if num < len(self.env_paths):
            self.env = self._make(self.env_paths[num])
        else:
            print("Wrong number,")

    # main methods
    def step(self, num_steps=-1, actions=None, min_dist=0.1):
        if num_steps < 1:
            num_steps = self.sim_step
        try:
            for _ in range(num_steps):
                self.env.step()
                if self.env.diverged:
                    self.env.diverged = False
                    raise MujocoException("simulation diverged (NaN or |x| > 1e10); data was reset")
        except MujocoException:
            self.reset()
        readings, contact = self.get_sensor_sensordata()
        out = np.empty(2, dtype=object)
        out[0], out[1] = readings, contact
        return out

    def reset(self):
        current_stiffness = self.set_new_stiffness()
        self.env.reset()
        self.env.forward()
        if self.sim_start > 0:
            self.step(self.sim_start)
        return current_stiffness

    def get_sensor_sensordata(self):
        data = self.env.data
        # literal restatement of ref: manenv.py:68-83 on the per-step touch mask: `fingers_left` aliases the
        # class-level list and is emptied for good the first time every finger has touched the object;
        # from then on the flag is simply "there is at least one contact".
        is_contact_between_fingers_and_object = False
        fingers_left = self.finger_names
        touch = self.env.touch
        if data.ncon > 0:
            for k, finger_name in enumerate(self._finger_names0):
                if touch & (1 << k) and finger_name in fingers_left:
                    fingers_left.remove(finger_name)
            if len(fingers_left) == 0:
                is_contact_between_fingers_and_object = True
        return np.copy(data.sensordata), is_contact_between_fingers_and_object

    def toggle_grip(self):
        if self.is_closing:
            self.loose_hand()
        else:
            self.close_hand()

    def close_hand(self):
        for i in range(2):
            self.env.data.ctrl[i] = -0.2
        self.is_closing = True

    def loose_hand(self):
        for i in range(2):
            self.env.data.ctrl[i] = 0.2
        self.is_closing = False

    def set_new_stiffness(self, range_min=300, range_max=1400):
        new_value = np.random.uniform(range_min, range_max)
        for i in self.joint_ids:
            if i < self.env.model.jnt_stiffness.shape[0]:
                self.env.model.jnt_stiffness[i] = new_value
        for i in self.tendon_ids:
            self.env.model.tendon_stiffness[i] = new_value
        return new_value

    def get_env(self):
        return self.env

    def render(self):
        pass  # no viewer (component 10)

    # specs
    @staticmethod
    def get_std_spec(args):
        return {
            "sim_start": args.sim_start,
            "sim_step": args.sim_step,
            "env_paths": args.mujoco_model_paths,
            "is_vis": args.vis
        }
