"""BASELINE.json configs[4], the part that exists: the refined composite (softbox_refined, 434 shell elements, volume-tendon
damper 20) with two-finger contact, 1 M worlds over the GPUs of one box, every world with its own stiffness / shell damping /
object offset.  One full squeeze episode per world, device-timed, max over ranks.  (The multi-finger half of configs[4] needs
the 4-finger gripper, SURVEY section 8 row f4a: not built.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/stress_1m.py [worlds_total]
"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

W_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
batched = importlib.import_module("soft-grip_b200.batched")
lo, hi = bench.shard_range(W_total, rank, world)
Wg = hi - lo


class A:
    seed = 0; randomise = "all"; fixed_stiffness = None; damping_range = (100.0, 200.0); offset_range = 0.05


env = batched.BatchedManEnv(os.path.join(ROOT, "tests", "golden", "softbox_refined.sgm"), Wg, device=dev, dtype=torch.float32, seed=0, world_offset=lo)
k, d, off = bench.world_params(A, np.arange(lo, hi), 0)
env.set_params(damping=d, tendon_damping=np.full(Wg, 20.0), object_offset=off)
env.rollout(schedule=batched.default_schedule(env.nu, n_settle=1, n_iter=1), stiffness=k)      # warm-up: module load, first launch
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
traj, kk, st = env.rollout(stiffness=k)
e1.record()
torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
flags = torch.tensor([float(((st & 1) != 0).sum()), float(((st & 10) != 0).sum()), float(torch.isfinite(traj).all())], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(flags, op=dist.ReduceOp.SUM)
if rank == 0:
    steps = 1401
    print(json.dumps({"stress": "BASELINE configs[4] (refined composite, two fingers)", "model": "softbox_refined", "worlds_total": W_total,
                      "n_gpus": world, "worlds_per_gpu": Wg, "seconds": ms.item() / 1e3, "world_steps_per_s": W_total * steps / (ms.item() / 1e3),
                      "worlds_diverged": int(flags[0].item()), "worlds_capacity_or_unsupported": int(flags[1].item()),
                      "trajectory_finite_ranks": int(flags[2].item()), "geometry": env.config(),
                      "trajectory_bytes_per_gpu": int(traj.numel() * 4)}))
if world > 1:
    dist.destroy_process_group()
