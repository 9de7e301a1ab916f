"""Dev tool: single-step parity of the CUDA path against the oracle along an oracle trajectory."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sg = importlib.import_module("soft-grip_b200")
from oracle import sgoracle as so
import torch

name = sys.argv[1] if len(sys.argv) > 1 else "softbox"
prec = sys.argv[2] if len(sys.argv) > 2 else "64"
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
batched = importlib.import_module("soft-grip_b200.batched")
mjcf = importlib.import_module("soft-grip_b200.mjcf")
blob = os.path.join(ROOT, "tests", "golden", name + ".sgm")
model = mjcf.load_blob(blob)
om = so.OracleModel(open(blob, "rb").read())
ow = so.OracleWorld(om)
ow.set_geom_mask(batched.geom_name_mask(model.names["geom"], "OBJ", ("g12", "g2")))
k = 700.0
ow.set_stiffness(k)
env = batched.BatchedManEnv(blob, 2, dtype=torch.float64 if prec == "64" else torch.float32)
env.set_new_stiffness(stiffness=[k, k])
env.set_debug_world(0)
print("info nv", env.info.nv, "levels", env.info.nlevels, "maxcon", env.info.maxcon, "smem32", env.info.smem_bytes32, "smem64", env.info.smem_bytes64)
ow.reset()
worst = {}
def rel(a, b):
    return np.abs(a - b).max() / max(1e-12, np.abs(b).max())
for t in range(nsteps):
    if t == 281: ow.set_ctrl([-0.2, -0.2]); ctrl = [-0.2, -0.2]
    if t == 0: ctrl = [0.0, 0.0]
    if t == 841: ow.set_ctrl([0.2, 0.2]); ctrl = [0.2, 0.2]
    q, v, a, w = ow.get_state()
    env.set_state(q, v, a, w)
    env.set_ctrl(np.tile(np.array(ctrl), (2, 1)))
    st = ow.step()
    sens, touch = env.step(1)
    q2, v2, a2, w2 = ow.get_state()
    gq, gv, ga, gw = env.get_state()
    e = dict(q=rel(gq[0], q2), v=rel(gv[0], v2), act=np.abs(ga[0] - a2).max(), qacc=rel(gw[0], w2),
             sens=rel(sens[0].double().cpu().numpy(), ow.sensordata()))
    ncon_o, nefc_o, it_o = ow.get_int("ncon"), ow.get_int("nefc"), ow.get_int("solver_iter")
    ncon_g, nefc_g, it_g = int(env.debug(0, "ncon")[0]), int(env.debug(0, "nefc")[0]), int(env.debug(0, "solver_iter")[0])
    bad = (ncon_o != ncon_g) or (nefc_o != nefc_g) or (it_o != it_g) or max(e.values()) > (1e-6 if prec == "64" else 1e-2)
    for kk, vv in e.items(): worst[kk] = max(worst.get(kk, 0), vv)
    if bad or t % 50 == 0:
        print(t, "ncon", ncon_o, ncon_g, "nefc", nefc_o, nefc_g, "iter", it_o, it_g, {kk: "%.2e" % vv for kk, vv in e.items()}, "touch", ow.touch_mask(), int(touch[0]) if False else "", "st", st, env.status()[0], "ncand", int(env.debug(0, "ncand")[0]))
    if bad and "-k" not in sys.argv:
        fo, fg = ow.get("efc_force"), env.debug(0, "efc_force")
        ao, ag = ow.get("efc_aref"), env.debug(0, "efc_aref")
        Ro, Rg = ow.get("efc_R"), env.debug(0, "efc_R")
        n = min(len(fo), len(fg))
        print(" efc_aref err", np.abs(ao[:n] - ag[:n]).max(), np.argmax(np.abs(ao[:n] - ag[:n])), " R err", np.abs(Ro[:n] - Rg[:n]).max(), np.argmax(np.abs(Ro[:n] - Rg[:n])), " force err", np.abs(fo[:n] - fg[:n]).max(), np.argmax(np.abs(fo[:n] - fg[:n])))
        do, dg = ow.get("con_dist"), env.debug(0, "con_dist")
        print(" con_dist", do[:8], dg[:8])
        po, pg = ow.get("con_pos"), env.debug(0, "con_pos")
        m = min(len(po), len(pg))
        if m: print(" con_pos err", np.abs(po[:m] - pg[:m]).max(), "frame err", np.abs(ow.get("con_frame")[:3*m] - env.debug(0, "con_frame")[:3*m]).max())
        fo_, fg_ = ow.get("con_frame").reshape(-1, 9), env.debug(0, "con_frame").reshape(-1, 9)
        po_, pg_ = po.reshape(-1, 3), pg.reshape(-1, 3)
        g1, g2 = ow.get("con_geom1"), ow.get("con_geom2")
        for ci in range(min(len(po_), len(pg_))):
            if np.abs(po_[ci] - pg_[ci]).max() > 1e-9 or np.abs(fo_[ci] - fg_[ci]).max() > 1e-9:
                print("  contact", ci, "geoms", g1[ci], g2[ci], "dist", do[ci], dg[ci], "\n    pos o", po_[ci], "g", pg_[ci], "\n    n o", fo_[ci][:3], "g", fg_[ci][:3], "\n    t1 o", fo_[ci][3:6], "g", fg_[ci][3:6])
        print(" qacc o", w2[:8], "\n qacc g", gw[0][:8])
        break
print("worst", {kk: "%.3e" % vv for kk, vv in worst.items()})
