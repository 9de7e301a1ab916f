"""TEST INFRASTRUCTURE: drives an AddressSanitizer build of the emulator library (kernel source + C-ABI host logic of
soft-grip_b200/csrc) through steps and rollouts -- precisions, lanes per world, multi-warp CTAs, the larger models -- so
that a heap out-of-bounds access in the kernel source or the host tables is caught on the CPU.
Run by tests/test_simt.py::test_emulated_step_kernel_is_clean_under_asan as
    LD_PRELOAD=libasan.so python asan_drive.py <repo root> <asan build of libsoftgrip_simt.so>"""
import importlib
import os
import sys

import numpy as np

root, libpath = sys.argv[1], sys.argv[2]
for p in (root, os.path.join(root, "tests"), os.path.join(root, "tests", "simt")):
    sys.path.insert(0, p)
import emu  # noqa: E402

emu.build = lambda: libpath
batched = importlib.import_module("soft-grip_b200.batched")
golden = os.path.join(root, "tests", "golden")
st = np.load(os.path.join(golden, "softbox_states.npz"))


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-12, np.abs(b).max()))


# (the last column forces the equality rows into tensor memory and the contact records through the shared-memory ring, or
# keeps both off; the driver's default decides otherwise)
for prec, lpw, nw, tmem in ((32, 8, None, "1"), (32, 8, None, "0"), (64, 8, 4, "1"), (32, 16, None, None), (64, 32, None, "1"), (32, 4, 2, None)):
    if tmem is None:
        os.environ.pop("SOFTGRIP_TMEM", None)
    else:
        os.environ["SOFTGRIP_TMEM"] = tmem
    W = 32 // lpw + 1 if nw is None else 19
    env = emu.EmuBatch(os.path.join(golden, "softbox.sgm"), W, prec=prec, lpw=lpw, nw=nw)
    env.set_params(stiffness=np.full(W, 700.0))
    env.set_debug_world(W - 1)
    for i in (0, 4, 7, 9, 11):                     # contact-free, first contacts, peak contact (58), release
        env.set_state(st["q"][i], st["v"][i], st["act"][i], st["warm"][i])
        env.set_ctrl([st["ctrl"][i]] * 2)
        env.step(1)
        assert rel(env.get_state()[0][-1], st["q1"][i]) < 1e-4, (prec, lpw, i)
    traj, touch, status = env.rollout(batched.default_schedule(2, n_settle=1, n_iter=3, open_close_div=2))
    assert np.isfinite(traj).all()
    env.close()
os.environ["SOFTGRIP_TMEM"] = "1"
for name, lpw, td in (("softbox_refined", 32, 20.0), ("softball", 8, 50.0), ("softcylinder", 16, 50.0)):
    env = emu.EmuBatch(os.path.join(golden, name + ".sgm"), 2, prec=32, lpw=lpw)
    env.set_params(stiffness=np.full(2, 700.0), tdamping=np.full(2, td))
    env.reset()
    env.set_ctrl([-0.2, -0.2])
    env.step(3)
    traj, touch, status = env.rollout(batched.default_schedule(2, n_settle=1, n_iter=2, open_close_div=2))
    assert np.isfinite(traj).all()
    env.close()
print("ASAN DRIVE DONE")
