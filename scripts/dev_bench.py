"""Dev tool: rollout timing + fp32/fp64 trace comparison."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
batched = importlib.import_module("soft-grip_b200.batched")
name = sys.argv[1] if len(sys.argv) > 1 else "softbox"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
blob = os.path.join(ROOT, "tests", "golden", name + ".sgm")
for prec in (torch.float32, torch.float64):
    Wp = W if prec == torch.float32 else max(256, W // 8)
    env = batched.BatchedManEnv(blob, Wp, dtype=prec, seed=0)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        traj, k, st = env.rollout()
        torch.cuda.synchronize(); dt = time.time() - t
        print(prec, "W", Wp, "rollout", dt, "s  world-steps/s", Wp * 1401 / dt, "status nonzero", int((st != 0).sum()), "nan", bool(torch.isnan(traj).any()))
    if prec == torch.float32: t32 = traj[:256].double().cpu().numpy(); k32 = k[:256].cpu().numpy()
    else: t64 = traj[:256].cpu().numpy(); k64 = k[:256].cpu().numpy()
assert np.allclose(k32, k64)
err = np.abs(t32 - t64)
scale = np.abs(t64).max(axis=(0, 1))
print("fp32 vs fp64 traj: max abs err per channel", err.max(axis=(0, 1)))
print("channel scale", scale)
print("rel (max over rows of err/scale)", (err.max(axis=(0, 1)) / scale))
print("median abs err", np.median(err, axis=(0, 1)))
