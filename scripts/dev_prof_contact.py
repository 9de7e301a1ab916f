"""Dev tool: a short contact-heavy launch for ncu.  Every world starts from the golden snapshot of physics step 845 of the
softbox squeeze episode (58 contacts, tests/golden/softbox_states.npz) and takes `nsub` physics steps through the step API,
so that the stall samples of one ncu capture cover contact-rich steps only (the rollout's first 280 steps are contact-free
and fill the sampling buffer).  usage: dev_prof_contact.py [W=9472] [nsub=21]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
batched = importlib.import_module("soft-grip_b200.batched")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 21
S = np.load(os.path.join(ROOT, "tests", "golden", "softbox_states.npz"))
i = list(S["step"]).index(845)
env = batched.BatchedManEnv(os.path.join(ROOT, "tests", "golden", "softbox.sgm"), W, dtype=torch.float32, seed=0)
env.set_new_stiffness(stiffness=[700.0] * W)
for rep in range(2):
    env.set_state(S["q"][i], S["v"][i], S["act"][i], S["warm"][i])
    env.set_ctrl(np.tile(np.asarray([S["ctrl"][i]] * 2, dtype=np.float64), (W, 1)))
    torch.cuda.synchronize(); t = time.time()
    env.step(nsub)
    torch.cuda.synchronize(); dt = time.time() - t
    print("W", W, "nsub", nsub, "time", dt, "world-steps/s", W * nsub / dt, "status!=0", int((env.status() != 0).sum()))
