"""Generates the golden vectors committed next to this script from the CPU oracle.

The reference ships no golden data for this path ("parity unpinned"), so these
vectors pin the *oracle* (regression guard across machines / compilers) and give the GPU tests inputs
that do not need /root/reference:

  softbox_episode_k700.npz   full create_dataset episode (200 rows x 12 channels, touch masks, final state)
  softbox_states.npz         states along that episode + the oracle's result of ONE more mj_step from each
  <model>_settle.npz         first 60 physics steps of every model (sensor rows, qpos norm)

    python tests/golden/make_golden.py          (needs only the committed .sgm blobs + oracle/)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import sgoracle as so  # noqa: E402

batched = importlib.import_module("soft-grip_b200.batched")
mjcf = importlib.import_module("soft-grip_b200.mjcf")

SNAP_STEPS = [0, 1, 140, 282, 300, 450, 600, 845, 850, 900, 1000, 1400]


def world(name, k=700.0):
    blob = open(os.path.join(HERE, name + ".sgm"), "rb").read()
    model = mjcf.load_blob(blob)
    om = so.OracleModel(blob)
    w = so.OracleWorld(om)
    w.set_geom_mask(batched.geom_name_mask(model.names["geom"], "OBJ", ("g12", "g2")))
    w.set_stiffness(k)
    return w


def ctrl_at(step):
    """ctrl in force while physics step `step` (0-based, after reset) executes, for the default protocol
    with sim_start=1, sim_step=7: close at env-step 40, loosen at env-step 120."""
    if step < 1 + 40 * 7:
        return 0.0
    if step < 1 + 120 * 7:
        return -0.2
    return 0.2


if __name__ == "__main__":
    w = world("softbox")
    rows, touch, st = w.episode()
    q, v, a, ws = w.get_state()
    np.savez_compressed(os.path.join(HERE, "softbox_episode_k700.npz"), rows=rows, touch=touch, status=st, qpos=q, qvel=v, act=a, warm=ws)
    print("episode status", st, "rows", rows.shape, "max|acc|", np.abs(rows[:, :6]).max())

    w = world("softbox")
    w.reset()
    snaps = {k: [] for k in ("step", "ctrl", "q", "v", "act", "warm", "q1", "v1", "act1", "qacc1", "sens1", "ncon1", "nefc1", "iter1", "touch1")}
    for step in range(max(SNAP_STEPS) + 1):
        c = ctrl_at(step)
        w.set_ctrl([c, c])
        if step in SNAP_STEPS:
            q, v, a, ws = w.get_state()
        w.step()
        if step in SNAP_STEPS:
            q1, v1, a1, ws1 = w.get_state()
            for key, val in (("step", step), ("ctrl", c), ("q", q), ("v", v), ("act", a), ("warm", ws), ("q1", q1), ("v1", v1),
                             ("act1", a1), ("qacc1", ws1), ("sens1", w.sensordata()), ("ncon1", w.get_int("ncon")),
                             ("nefc1", w.get_int("nefc")), ("iter1", w.get_int("solver_iter")), ("touch1", w.touch_mask())):
                snaps[key].append(val)
    np.savez_compressed(os.path.join(HERE, "softbox_states.npz"), **{k: np.array(vv) for k, vv in snaps.items()})
    print("snapshots", snaps["step"], "ncon", snaps["ncon1"])

    for name in ("softbox", "softball", "softcylinder"):
        w = world(name)
        w.reset()
        sens, qn, ncon = [], [], []
        for step in range(60):
            w.step()
            sens.append(w.sensordata())
            qn.append(np.abs(w.get_state()[0]).max())
            ncon.append(w.get_int("ncon"))
        np.savez_compressed(os.path.join(HERE, name + "_settle.npz"), sens=np.array(sens), qmax=np.array(qn), ncon=np.array(ncon), status=w.status())
        print(name, "settle status", w.status(), "ncon[0]", ncon[0], "qmax end", qn[-1])
