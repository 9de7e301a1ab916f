"""TEST INFRASTRUCTURE (oracle): numpy restatement of the two trajectory operations of the reference's trainer.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

* ``noised_modality``  follows ref: functions/optimization.py:6-14 (acc channels += N(0, 0.7), gyro += N(0, 0.06)).
  The reference draws from TensorFlow's unseeded global generator, so its *values* cannot be reproduced; what is
  restated bit-for-bit is the generator the CUDA path documents: Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
  "Parallel random numbers: as easy as 1, 2, 3", SC'11 -- the algorithm behind tf.random / curand / torch.cuda),
  pinned here by the published known-answer vectors of the Random123 distribution (``PHILOX_KAT``), followed by
  Box-Muller evaluated in float64.  Parity of the reference's semantics is distributional (mean 0, the two sigmas,
  which channels get which) and is tested as such.
* ``channel_mean_std`` is the reference's own numpy expression, ref: functions/utils.py:39-40.
* ``mask_contact`` follows ref: create_dataset.py:43-44,57-58 with the contact flag of ref: manenv.py:65-83, written the
  way the reference writes it (a finger-name list that contacts remove entries from), per world and per recorded row.
"""
import numpy as np

# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key, expected)
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32-valued array, key: (k0, k1) -> (..., 4) uint32."""
    c = np.asarray(counter, dtype=np.uint64) & _LO
    x, y, z, w = (c[..., i].copy() for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * x, _M1 * z
        x, y, z, w = (p1 >> np.uint64(32)) ^ y ^ np.uint64(k0), p1 & _LO, (p0 >> np.uint64(32)) ^ w ^ np.uint64(k1), p0 & _LO
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack([x, y, z, w], axis=-1).astype(np.uint32)


def _u01(x):
    """(0, 1] uniform of csrc/sg_traj.cuh: fp32(x) * 2^-32 + 2^-33 (the product is exact, the sum rounds once)."""
    return (x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)


def standard_normals(nelem, seed, first_elem=0):
    """The N(0,1) draw of every flat element index in [first_elem, first_elem + nelem), both multiples of 4: float64
    Box-Muller of the fp32 uniforms."""
    nq = nelem // 4
    q = np.arange(nq, dtype=np.uint64) + np.uint64(first_elem // 4)
    ctr = np.stack([q & _LO, q >> np.uint64(32), np.zeros_like(q), np.zeros_like(q)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    z = np.empty((nq, 4), dtype=np.float64)
    for a, b, o in ((0, 1, 0), (2, 3, 2)):
        rad = np.sqrt(-2.0 * np.log(_u01(r[:, a]).astype(np.float64)))
        ang = 2.0 * np.pi * _u01(r[:, b]).astype(np.float64)
        z[:, o], z[:, o + 1] = rad * np.cos(ang), rad * np.sin(ang)
    return z.reshape(-1)


def noised_modality(data, seed, sigma_acc=0.7, sigma_gyro=0.06, nacc=None, mean=None, std=None, first_row=0):
    data = np.asarray(data)
    nchan = data.shape[-1]
    nacc = nchan // 2 if nacc is None else nacc
    sig = np.where(np.arange(nchan) < nacc, np.float32(sigma_acc), np.float32(sigma_gyro)).astype(np.float64)
    z = standard_normals(data.size, seed, first_elem=first_row * nchan).reshape(data.shape)
    out = data.astype(np.float64) + sig * z
    if mean is not None:
        out = (out - np.asarray(mean, dtype=np.float64).reshape(-1)) / np.asarray(std, dtype=np.float64).reshape(-1)
    return out


def channel_mean_std(train_x):
    train_x = np.asarray(train_x, dtype=np.float64)
    ax = tuple(range(train_x.ndim - 1))
    return np.mean(train_x, axis=ax, keepdims=True), np.std(train_x, axis=ax, keepdims=True)


def contact_flag_literal(fingers_left, row_fingers, ncon_positive):
    """One call of get_sensor_sensordata (ref: manenv.py:65-83) reduced to what decides its flag.  ``fingers_left`` is the
    class-level list the method aliases and mutates (ref: manenv.py:70,80); ``row_fingers`` are the finger names that
    touch an object geom in this row's contacts.  While the list is non-empty the flag turns true when the row's contacts
    empty it; once it is empty, ``len(fingers_left) == 0`` holds at the first contact, so the flag is ``ncon >= 1``."""
    if len(fingers_left) == 0:
        return bool(ncon_positive)
    for name in list(row_fingers):
        if name in fingers_left:
            fingers_left.remove(name)
    return len(fingers_left) == 0


def mask_contact(traj, touch, nfingers, any_bit, mode, fingers_left=None):
    """traj (W,T,C), touch (W,T) int bit masks (bit k: finger k touches an object geom in that row; any_bit: ncon >= 1).
    mode "intended": flag = every finger touches in this row.  mode "reference": the literal list semantics, one list per
    world, carried across rows (and across calls when ``fingers_left`` -- a list of per-world lists -- is given).
    Returns the masked copy (rows with a false flag zeroed, ref: create_dataset.py:43-44,57-58)."""
    traj = np.array(traj, copy=True)
    touch = np.asarray(touch)
    W, T = touch.shape
    for w in range(W):
        left = list(range(nfingers)) if fingers_left is None else fingers_left[w]
        for t in range(T):
            row = [k for k in range(nfingers) if (int(touch[w, t]) >> k) & 1]
            if mode == "intended":
                flag = len(row) == nfingers
            else:
                flag = contact_flag_literal(left, row, int(touch[w, t]) & any_bit)
            if not flag:
                traj[w, t] = 0
    return traj
