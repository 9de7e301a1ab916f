"""Dev tool: short rollout for ncu."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
batched = importlib.import_module("soft-grip_b200.batched")
name = sys.argv[1] if len(sys.argv) > 1 else "softbox"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1776
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 60
blob = os.path.join(ROOT, "tests", "golden", name + ".sgm")
env = batched.BatchedManEnv(blob, W, dtype=torch.float32, seed=0)
nset = int(os.environ.get('PROF_SETTLE', '10'))
ev, val = batched.default_schedule(2) if rows == 200 else batched.default_schedule(2, n_settle=nset, n_iter=rows - nset, open_close_div=80)
for rep in range(int(os.environ.get('PROF_REPS', '2'))):
    torch.cuda.synchronize(); t = time.time()
    traj, k, st = env.rollout(schedule=(ev, val))
    torch.cuda.synchronize(); dt = time.time() - t
    print("W", W, "rows", rows, "time", dt, "world-steps/s", W * (1 + 7 * rows) / dt)
