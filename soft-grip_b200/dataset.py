"""Dataset writer/reader for the squeeze-episode sensor traces.

File format of the reference (ref: create_dataset.py:75-78, read at functions/utils.py:8-23):
    pickle of {"data": list of N float64 arrays (200, 12), "stiffness": list of N floats}
with channel order [acc(sensor1) 3, acc(sensor2) 3, gyro(sensor1) 3, gyro(sensor2) 3]
(ref: data/gripper/soft_grip_two_fingers.xml:118-123; the consumer splits [:, :, :6] / [:, :, 6:],
ref: functions/optimization.py:8).  A tensor variant (one .npz) is offered for large regenerations.
"""
import os
import pickle

import numpy as np


def to_reference_dict(traj, stiffness):
    """traj: (N, T, 12) array-like, stiffness: (N,) -> the reference's pickle payload."""
    traj = np.asarray(traj, dtype=np.float64)
    stiffness = np.asarray(stiffness, dtype=np.float64)
    if traj.ndim != 3 or traj.shape[0] != stiffness.shape[0]:
        raise ValueError("traj must be (N, T, C) and stiffness (N,)")
    return {"data": [np.array(traj[i]) for i in range(traj.shape[0])],
            "stiffness": [float(k) for k in stiffness]}


def write_pickle(path, traj, stiffness):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(to_reference_dict(traj, stiffness), f)


def read_pickle(path):
    """What functions/utils.create_tf_generators does with a dataset file: np.array(ds["data"]) -> (N,T,12)."""
    with open(path, "rb") as f:
        ds = pickle.load(f)
    return np.array(ds["data"]), np.array(ds["stiffness"])


def write_npz(path, traj, stiffness, status=None):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    extra = {} if status is None else {"status": np.asarray(status)}
    np.savez(path, data=np.asarray(traj), stiffness=np.asarray(stiffness), **extra)


def mask_contact(traj, contact):
    """--mask-contact of the reference: rows without finger-object contact are zeroed
    (ref: create_dataset.py:43-44,57-58)."""
    traj = np.array(traj, copy=True)
    traj[~np.asarray(contact, dtype=bool)] = 0
    return traj


def feature_stats(traj, stiffness, nbins=4, lo=300.0, hi=1400.0):
    """Per-stiffness-bin statistics of the |acc| and |gyro| magnitudes per sensor (mean / std / peak),
    the comparison SURVEY section 8d cfg 4 asks for instead of per-class statistics."""
    traj = np.asarray(traj, dtype=np.float64)
    k = np.asarray(stiffness, dtype=np.float64)
    mags = np.stack([np.linalg.norm(traj[..., 3 * c:3 * c + 3], axis=-1) for c in range(traj.shape[-1] // 3)], axis=-1)
    edges = np.linspace(lo, hi, nbins + 1)
    out = []
    for b in range(nbins):
        sel = (k >= edges[b]) & (k < edges[b + 1] if b < nbins - 1 else k <= edges[b + 1])
        if not sel.any():
            out.append(None)
            continue
        m = mags[sel]
        out.append({"n": int(sel.sum()), "mean": m.mean(axis=(0, 1)), "std": m.std(axis=(0, 1)), "peak": m.max(axis=1).mean(axis=0)})
    return edges, out
