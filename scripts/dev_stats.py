"""Dev helper: contact-count / solver-iteration / GS-level statistics of an episode (oracle)."""
import sys, os, numpy as np, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sgoracle as so
mjcf = importlib.import_module("soft-grip_b200.mjcf")
name = sys.argv[1] if len(sys.argv) > 1 else "softbox"
ks = [float(x) for x in sys.argv[2:]] or [300., 700., 1400.]
blob = open("tests/golden/%s.sgm" % name, "rb").read()
om = so.OracleModel(blob)
for k in ks:
    w = so.OracleWorld(om); w.set_stiffness(k); w.reset(); w.forward()
    ncon, nrows, it, st = [], [], [], 0
    ctrl = np.zeros(2)
    for t in range(1401):
        envstep = (t - 1) // 7
        if t == 1 + 40 * 7: ctrl[:] = -0.2; w.set_ctrl(ctrl)
        if t == 1 + 120 * 7: ctrl[:] = 0.2; w.set_ctrl(ctrl)
        st |= w.step()
        ncon.append(w.get_int("ncon")); it.append(w.get_int("solver_iter")); nrows.append(w.get_int("nefc"))
    ncon = np.array(ncon); it = np.array(it); nrows = np.array(nrows)
    print(name, "k=%g" % k, "status", st, "ncon max/mean", ncon.max(), ncon.mean(), "nefc max", nrows.max(), "iters mean/min", it.mean(), it.min(),
          "ncon by phase", ncon[:281].max(), ncon[281:841].max(), ncon[841:].max())
