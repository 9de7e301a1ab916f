#!/usr/bin/env python
"""Benchmark of the soft-grip hot path on B200: world-steps/s of the batched squeeze-episode rollout.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels behind the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU path (oracle port, all host cores)
    N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full squeeze episode (reset + 1 + 200*7 = 1401 physics steps, ref: create_dataset.py:14-17,
33-60) of a batch of `--worlds-per-gpu` worlds per GPU, i.e. one launch of the rollout kernel per GPU.
Workload = BASELINE.json configs[2] (randomised stiffness U(300,1400), fp32 fast path, worlds sharded over
the GPUs with no collective on the step path), on the softbox model (the stable primary model),
processed in per-GPU batches so that one step takes seconds, not minutes.  Weak scaling: per-GPU work fixed.

`value`   : device-timed (CUDA events on the launching stream, max over ranks), parameters resident in HBM.
`e2e`     : the same episode through sg_batch_rollout_host with pinned HOST buffers (H2D of the stiffness
            vector and D2H of the whole trajectory inside the timed region, wall clock, max over ranks).
`roofline`: HBM roofline of the rollout kernel on its algorithmic bytes (the kernel is latency/issue bound on
            the Gauss-Seidel sweep, so this fraction is tiny by construction).
`cpu_baseline`: the fp64 oracle (a port, not libmujoco) on the host cores, bounded sample, rank 0 at N=1.
`sensor_trace_error`: (N=1) four of the worlds just simulated against the fp64 oracle with the same stiffness; informational.
`traj_kernels`: (N=1, outside the timed region) the HBM-bound kernels that post-process the trajectory buffer, each timed
            alone against the HBM roof; informational, a failure there is reported in the key and never raised.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STEPS_PER_EPISODE = 1 + 200 * 7
METRIC = "world-steps/sec"


# ---------------------------------------------------------------------------------------------
# helpers shared with the tests (pure host logic)
# ---------------------------------------------------------------------------------------------
def shard_range(total, rank, world):
    """Contiguous block of worlds owned by `rank`."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def max_over_ranks(x, device):
    import torch
    d = _dist()
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device):
    import torch
    d = _dist()
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------
# CPU path (oracle port): used as cpu_baseline and as the --impl reference arm
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(blob_path, tendon_damping=None):
    from oracle import sgoracle as so
    blob = open(blob_path, "rb").read()
    _W["m"] = so.OracleModel(blob)
    _W["w"] = so.OracleWorld(_W["m"])
    if tendon_damping is not None:
        _W["w"].set_tendon_damping(0, float(tendon_damping))


def _cpu_episodes(ks):
    w = _W["w"]
    chk, flops, n = 0.0, 0.0, 0
    for k in ks:
        w.set_stiffness(float(k))
        rows, touch, st = w.episode()
        chk += float(np.abs(rows).sum())
        flops += w.last_step_flops()
        n += 1
    return chk, flops / max(1, n)


def cpu_throughput(blob_path, episodes_per_core, repeats=1, cores=None, seed=0, tendon_damping=None, model="softbox"):
    """world-steps/s of the oracle with one process per host core; returns (value, cores, seconds, sample text)."""
    from concurrent.futures import ProcessPoolExecutor
    batched = importlib.import_module("soft-grip_b200.batched")
    cores = cores or len(os.sched_getaffinity(0))
    best = None
    with ProcessPoolExecutor(cores, initializer=_cpu_init, initargs=(blob_path, tendon_damping)) as ex:
        list(ex.map(_cpu_episodes, [[700.0][:0]] * cores))      # spin the workers up (model load excluded)
        for rep in range(repeats):
            ks = batched.world_uniform(seed + rep, np.arange(cores * episodes_per_core), 300, 1400)
            chunks = [list(ks[i::cores]) for i in range(cores)]
            t0 = time.perf_counter()
            res = list(ex.map(_cpu_episodes, chunks))
            dt = time.perf_counter() - t0
            val = cores * episodes_per_core * STEPS_PER_EPISODE / dt
            if best is None or val > best[0]:
                best = (val, dt, res[0][1])
    sample = "%d processes x %d full squeeze episodes (%s, stiffness U(300,1400), fp64 oracle port)" % (cores, episodes_per_core, model)
    return best[0], cores, best[1], sample, best[2]


def sensor_trace_error(blob_path, rows, ks, tendon_damping=None):
    """The metric's second half ("sensor-trace error"), informational: the sensor traces the bench just produced for a few
    worlds against the fp64 oracle run on the host with the same stiffness (the oracle as the checker, inside the
    cpu_baseline leg).  Relative to each channel's peak; the settle rows are contact-free and must agree tightly, over the
    squeeze contact make/break events amplify round-off, so the median row is reported (tests/test_gpu.py holds the bars)."""
    try:
        _cpu_init(blob_path, tendon_damping)
        w = _W["w"]
        settle, med, worst = 0.0, [], 0.0
        for r, k in zip(rows, ks):
            w.set_stiffness(float(k))
            want, _, st = w.episode()
            scale = np.abs(want).max(axis=0) + 1e-12
            err = (np.abs(np.asarray(r, dtype=np.float64) - want) / scale).max(axis=1)
            settle = max(settle, float(err[:40].max()))
            med.append(float(np.median(err)))
            worst = max(worst, float(err.max()))
        return {"worlds": len(med), "settle_rows_max_rel": settle, "row_median_rel": float(np.median(med)), "row_max_rel": worst,
                "against": "fp64 oracle port on the host (not libmujoco), same stiffness, relative to each channel's peak",
                "stated_fp32_tolerance": {"settle_rows": 1e-4, "row_median": 5e-3}}
    except Exception as e:                                   # noqa: BLE001 -- informational key only
        return {"error": "%s: %s" % (type(e).__name__, e)}


# ---------------------------------------------------------------------------------------------
def run_reference_arm(args, rank):
    if rank != 0:
        return
    blob = os.path.join(ROOT, "tests", "golden", args.model + ".sgm")
    from oracle import sgoracle as so
    so.build()
    cores = len(os.sched_getaffinity(0))
    from concurrent.futures import ProcessPoolExecutor
    batched = importlib.import_module("soft-grip_b200.batched")
    times = []
    with ProcessPoolExecutor(cores, initializer=_cpu_init, initargs=(blob, args.tendon_damping)) as ex:
        list(ex.map(_cpu_episodes, [[]] * cores))
        for it in range(args.warmup + args.steps):
            ks = batched.world_uniform(args.seed + it, np.arange(cores * args.ref_episodes_per_core), 300, 1400)
            t0 = time.perf_counter()
            list(ex.map(_cpu_episodes, [list(ks[i::cores]) for i in range(cores)]))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = args.steps * cores * args.ref_episodes_per_core * STEPS_PER_EPISODE / total
    sample = "%d processes x %d episodes per step" % (cores, args.ref_episodes_per_core)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "world-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, per_gpu=cores * args.ref_episodes_per_core, note="CPU path: fp64 oracle port of the MuJoCo step (not libmujoco), host cores only"),
            "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu, note=None):
    cfg = {"workload": "BASELINE.json configs[2]: %s model, randomised stiffness U(300,1400) on joints 11..63 + tendon 0, fp32 fast path, "
                       "full squeeze horizon (1401 physics steps, 200 sensor rows x 12 channels per world), worlds sharded over the GPUs" % args.model,
           "model": args.model, "worlds_per_gpu_per_step": per_gpu, "physics_steps_per_world_per_step": STEPS_PER_EPISODE,
           "sim_step": 7, "sim_start": 1, "rows": 200, "parallelism": "independent world shards, no collective on the step path",
           "l2": "256 MiB scratch buffer written between timed steps (L2 flush); kernel inputs are tiny, state lives in shared memory"}
    if getattr(args, "tendon_damping", None) is not None:
        cfg["tendon_damping"] = ("volume-tendon damper set to %g (model file: 100): the value at which this model is stable under the "
                                 "restated engine semantics, SURVEY App. E / section 8d cfg 3" % args.tendon_damping)
    if note:
        cfg["note"] = note
    return cfg


def measure_traj_kernels(torch, traj, flush, hbm_peak, reps=5):
    """Outside the timed region, N = 1 only: the HBM-bound kernels that post-process the trajectory buffer where the rollout
    left it (csrc/sg_traj.cuh: noise augmentation, channel statistics), each timed alone with CUDA events on the current
    stream after an L2 flush; achieved = algorithmic bytes (noise: read + write, statistics: read) / average launch time.
    Extra information next to the headline numbers: a failure here is reported in the key, never raised."""
    try:
        fn = importlib.import_module("soft-grip_b200.functions")
        out = torch.empty_like(traj)
        nbytes = traj.numel() * traj.element_size()
        cases = (("sg_traj_noise_kernel<float>", lambda: fn.noised_modality(traj, seed=1, out=out), 2 * nbytes),
                 ("sg_traj_stats_partial_kernel<float> + final", lambda: fn.channel_mean_std(traj), nbytes))
        res = {}
        for name, f, algo in cases:
            ms = []
            for r in range(reps + 2):
                flush.fill_(r & 255)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                f()
                e1.record()
                torch.cuda.synchronize()
                if r >= 2:
                    ms.append(e0.elapsed_time(e1))
            avg = sum(ms) / len(ms)
            gbs = algo / (avg * 1e-3) / 1e9
            res[name] = {"ms": avg, "algorithmic_bytes": algo, "achieved_GBps": gbs, "peak_GBps": hbm_peak, "frac": gbs / hbm_peak}
        mean, std = fn.channel_mean_std(out)
        ref_std = (out - traj).double().std(dim=(0, 1), unbiased=False)
        res["check"] = {"noise_std_acc": float(ref_std[:6].mean()), "noise_std_gyro": float(ref_std[6:].mean()),
                        "stats_vs_torch_max_rel": float(((mean.reshape(-1) - out.double().mean(dim=(0, 1))).abs()
                                                         / (out.double().std(dim=(0, 1)) + 1e-30)).max())}
        return res
    except Exception as e:                                   # noqa: BLE001 -- informational key only
        return {"error": "%s: %s" % (type(e).__name__, e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="softgrip", choices=["softgrip", "reference"])
    ap.add_argument("--model", default="softbox")
    ap.add_argument("--worlds-per-gpu", type=int, default=18944)   # 2 x (148 SMs x 64 resident worlds)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tendon-damping", type=float, default=None,
                    help="override the composite's volume-tendon damper for every world (softball / softcylinder: 50 is stable)")
    ap.add_argument("--cpu-episodes-per-core", type=int, default=24)
    ap.add_argument("--ref-episodes-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                      # timing rules: W >= 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    batched = importlib.import_module("soft-grip_b200.batched")
    from importlib import import_module
    lib = import_module("soft-grip_b200._lib")

    Wg = args.worlds_per_gpu
    blob = os.path.join(ROOT, "tests", "golden", args.model + ".sgm")
    env = batched.BatchedManEnv(blob, Wg, device=dev, dtype=torch.float32, seed=args.seed, world_offset=rank * Wg)
    if args.tendon_damping is not None:
        env.set_params(tendon_damping=np.full(Wg, args.tendon_damping))
    ev, val = batched.default_schedule(env.nu)
    T = ev.shape[0]
    sc = lib.SgSchedule(1, 7, T, ev.ctypes.data_as(C.POINTER(C.c_int)), val.ctypes.data_as(C.POINTER(C.c_double)))
    traj = torch.empty((Wg, T, env.nsd), dtype=torch.float32, device=dev)
    ids = np.arange(rank * Wg, (rank + 1) * Wg)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def stiffness_for(it):
        return torch.from_numpy(batched.world_uniform(args.seed, ids, 300, 1400, stream=it)).to(dev)

    def one_step(it, timed):
        k = stiffness_for(it)                           # resident in HBM before the timed region
        env.stiffness = k
        env._push_params()
        flush.fill_(it & 255)                           # L2 flush
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        lib.check(env.L.sg_batch_rollout(env.h, C.byref(sc), C.c_void_p(traj.data_ptr()), None, C.c_void_p(stream.cuda_stream)))
        if timed:
            e1.record(stream)
            return e0, e1
        return None

    for it in range(args.warmup):
        one_step(it, False)
    torch.cuda.synchronize(dev)
    barrier()
    launches0 = env.launch_count()
    with ClockSampler(local_rank) as clk:
        evs = [one_step(args.warmup + it, True) for it in range(args.steps)]
        torch.cuda.synchronize(dev)
    barrier()
    launches = env.launch_count() - launches0
    ms_local = sum(a.elapsed_time(b) for a, b in evs)
    ms_total = max_over_ranks(ms_local, dev)
    st = env.status(clear=True)
    ndiv = sum_over_ranks(float(((st & batched.ST_DIVERGED) != 0).sum()), dev)
    nfull = sum_over_ranks(float(((st & (batched.ST_CON_FULL | batched.ST_UNSUPPORTED)) != 0).sum()), dev)
    finite = bool(torch.isfinite(traj).all().item())
    world_steps = float(world) * Wg * STEPS_PER_EPISODE * args.steps
    value = world_steps / (ms_total * 1e-3)

    # ---- end to end through the C-ABI with pinned HOST buffers ----
    k_host = torch.empty(Wg, dtype=torch.float64).pin_memory()
    traj_host = torch.empty((Wg, T, env.nsd), dtype=torch.float32).pin_memory()
    status_host = np.zeros(Wg, dtype=np.int32)

    def e2e_step(it):
        k_host.copy_(torch.from_numpy(batched.world_uniform(args.seed, ids, 300, 1400, stream=1000 + it)))
        lib.check(env.L.sg_batch_rollout_host(env.h, C.byref(sc), C.c_void_p(k_host.data_ptr()), C.c_void_p(traj_host.data_ptr()), None,
                                              status_host.ctypes.data_as(C.c_void_p)))
    e2e_step(0)
    torch.cuda.synchronize(dev)
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        e2e_step(1 + it)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    barrier()
    e2e_value = world_steps / e2e_s

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    # ---- roofline of the rollout kernel (HBM; algorithmic bytes = stiffness in + trajectory + status out) ----
    peaks, which = measured_peaks()
    bytes_per_world = 8 + T * env.nsd * 4 + 4
    launch_s = (ms_local * 1e-3) / args.steps
    achieved = Wg * bytes_per_world / launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else "fallback 6650 GB/s",
                "kernel": "sg_step_kernel2<float, 8> (rollout mode)", "algorithmic_bytes_per_world_episode": bytes_per_world,
                "note": "latency/issue-bound Gauss-Seidel kernel: state stays in shared memory for the whole episode, so the HBM fraction is tiny by design"}

    cpu = None
    pgs_flops = None
    trace_err = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import sgoracle as so
        so.build()
        last = args.warmup + args.steps - 1                  # `traj` still holds the last timed step
        k_last = batched.world_uniform(args.seed, ids, 300, 1400, stream=last)
        pick = [0, Wg // 3, (2 * Wg) // 3, Wg - 1]
        try:
            trace_err = sensor_trace_error(blob, traj[pick].double().cpu().numpy(), k_last[pick], args.tendon_damping)
        except Exception as e:                               # noqa: BLE001 -- informational key only
            trace_err = {"error": "%s: %s" % (type(e).__name__, e)}
        v, cores, secs, sample, pgs_flops = cpu_throughput(blob, args.cpu_episodes_per_core, seed=args.seed, tendon_damping=args.tendon_damping, model=args.model)
        cpu = {"value": v, "unit": "world-steps/s", "cores": cores, "kind": "port", "sample": sample + ", %.1f s" % secs}

    traj_kernels = None
    if world == 1:
        traj_kernels = measure_traj_kernels(torch, traj, flush, peaks["hbm_gbs"])

    line = {"metric": METRIC, "value": value, "unit": "world-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, Wg),
            "e2e": {"value": e2e_value, "unit": "world-steps/s", "h2d_bytes_per_step": Wg * 8, "d2h_bytes_per_step": Wg * (T * env.nsd * 4 + 4)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
            "launch_geometry": env.config(), "worlds_diverged": int(ndiv), "worlds_capacity_or_unsupported": int(nfull), "trajectory_finite": finite}
    if pgs_flops:
        fp32_peak = 148 * 128 * 2 * (line["clocks"]["sm_mhz"] or 1965.0) * 1e6 / 1e12
        line["fp32_pipe"] = {"pgs_flops_per_world_step_last": pgs_flops, "achieved_tflops": value / world * pgs_flops / 1e12,
                             "peak_tflops_at_sampled_clock": fp32_peak}
    if traj_kernels is not None:
        line["traj_kernels"] = traj_kernels
    if trace_err is not None:
        line["sensor_trace_error"] = trace_err
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
