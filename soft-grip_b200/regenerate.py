"""Regenerate the simulated part of the paper's dataset tree in one go (BASELINE.json configs[3]).

The reference produces these files by running ``create_dataset.py`` once per file with ``NUM_EPISODES`` edited by hand
(ref: create_dataset.py:14,83-93) and consumes them in ``run_experiments.sh`` (ref: run_experiments.sh:2-7):

    sim_box/train.pickle  sim_box/val.pickle                                   (stage 1 and 3: softbox only)
    sim_all/train.pickle  sim_all/{softball,softbox,softcylinder}_testing.pickle   (stage 2: all three shapes)

every file in the layout of ref: create_dataset.py:75-78 (``{"data": [N x (200,12) f64], "stiffness": [N x f64]}``).
The repo does not state the paper's sample counts (SURVEY section 8d cfg 4), so they are parameters.

Each sample is one squeeze episode of one world; samples of a model are numbered by a *global world id* and the
per-world stiffness is drawn from that id (``batched.world_uniform``), so the files are disjoint by construction and do
not depend on how many worlds went into a launch or on how many GPUs were used.  Per file a ``*.stats.npz`` holds the
per-channel mean / std computed on the device (ref: functions/utils.py:39-40) and the per-stiffness-bin feature
statistics; ``--npz`` also writes the tensors ``(N,200,12)`` / ``(N,)`` for loaders that do not want Python lists.

    python regenerate.py --out data/experiments --train 4096 --val 512 --test 512 \
        --softbox A.xml --softball B.xml --softcylinder C.xml
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 regenerate.py ...   # 8 GPUs

Multi-GPU: every rank simulates a contiguous slice of each file's world ids and writes a shard file; after one barrier
rank 0 concatenates the shards in rank order (the "final gather of dataset shards" -- the only exchange; nothing is
communicated on the step path) and computes the per-file statistics on the merged samples.  The merged pickles are
identical to a single-GPU run, `--noise-seed` included (a sample's draws depend on its global world id only).
"""
import importlib
import os
import sys
import time
from argparse import ArgumentParser

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))
_PKG = os.path.basename(_HERE)

SHAPES = ("softball", "softbox", "softcylinder")      # order of the stage-2 command line (ref: run_experiments.sh:7)


def plan_files(n_train, n_val, n_test, shapes=SHAPES):
    """-> list of (relative file stem, [(shape, first global world id, count), ...]).

    World ids of a shape are handed out consecutively, file after file, so no two files share a sample."""
    nxt = {s: 0 for s in shapes}

    def take(shape, n):
        part = (shape, nxt[shape], int(n))
        nxt[shape] += int(n)
        return part

    files = []
    if "softbox" in shapes:
        files.append(("sim_box/train", [take("softbox", n_train)]))
        files.append(("sim_box/val", [take("softbox", n_val)]))
    files.append(("sim_all/train", [take(s, n_train // len(shapes) + (i < n_train % len(shapes))) for i, s in enumerate(shapes)]))
    for s in shapes:
        files.append(("sim_all/{}_testing".format(s), [take(s, n_test)]))
    return [(stem, [p for p in parts if p[2] > 0]) for stem, parts in files]


class DeviceRollouts:
    """Episodes of global world ids [first, first + count) of one model, ``worlds_per_launch`` at a time.

    Yields device tensors (traj [n,200,12] fp32, stiffness [n] fp64, status [n] int32); masking happens on the device."""

    def __init__(self, model_paths, seed=0, worlds_per_launch=16384, device="cuda:0", mask_contact=False, contact_mode="intended",
                 sim_start=1, sim_step=7, tendon_damping=None):
        """tendon_damping: {shape: value} overrides the composite's volume-tendon damper of that shape for every world
        (softball / softcylinder run away at the committed 100 under the restated semantics; 50 is
        the value they are parity-tested at)."""
        self.batched = importlib.import_module(_PKG + ".batched")
        self.paths, self.seed, self.wpl, self.device = dict(model_paths), int(seed), int(worlds_per_launch), device
        self.mask, self.mode, self.sim_start, self.sim_step = bool(mask_contact), contact_mode, sim_start, sim_step
        self.tdamp = dict(tendon_damping or {})
        self._dm = {}
        self.seconds = {"model": 0.0, "simulate": 0.0}      # wall time spent compiling / uploading models and in rollouts
        self.world_steps = 0

    def shape_seed(self, shape):
        """The stiffness draw of a sample depends on (seed, shape, global world id): every shape has its own stream, so
        `softball_testing` and `softcylinder_testing` do not share a label vector."""
        return self.seed * len(SHAPES) + SHAPES.index(shape)

    def prefetch(self, files):
        """Simulate every world a file plan needs, one shape at a time and in as few launches as `worlds_per_launch`
        allows.  A launch is bound by the latency of one episode (about a second whether it holds 500 worlds or 9 000), so
        six files of a few hundred to a few thousand samples cost three launches instead of nine.  `files` is the output
        of plan_files; world ids of a shape are handed out consecutively, so [0, total) covers all its parts."""
        import torch
        total = {}
        for _, parts in files:
            for shape, first, count in parts:
                total[shape] = max(total.get(shape, 0), first + count)
        self._cache = {}
        for shape, n in total.items():
            chunks = list(self._simulate(shape, 0, n))
            self._cache[shape] = tuple(torch.cat([c[i] for c in chunks], 0) for i in range(3))

    def __call__(self, shape, first, count, noise_seed=None):
        """noise_seed: also apply the trainer's noise augmentation (ref: functions/optimization.py:6-14) on the device; the
        draw of a sample depends only on (noise_seed, shape, global world id), not on launches or ranks."""
        cache = getattr(self, "_cache", {}).get(shape)
        if cache is not None and first + count <= cache[0].shape[0]:
            traj, k, st = (t[first:first + count] for t in cache)
            if noise_seed is not None:
                traj = self._noise(traj.clone(), shape, first, noise_seed)
            yield traj, k, st
            return
        for traj, k, st in self._simulate(shape, first, count):
            if noise_seed is not None:
                traj = self._noise(traj, shape, first, noise_seed)
                first += int(traj.shape[0])
            yield traj, k, st

    def _noise(self, traj, shape, first, noise_seed):
        fn = importlib.import_module(_PKG + ".functions")
        fn.noised_modality(traj, seed=int(noise_seed) * len(SHAPES) + SHAPES.index(shape), out=traj, first_row=first * int(traj.shape[1]))
        return traj

    def _simulate(self, shape, first, count):
        import torch
        b = self.batched
        if shape not in self._dm:
            t0 = time.perf_counter()
            self._dm[shape] = b.DeviceModel(self.paths[shape])
            self.seconds["model"] += time.perf_counter() - t0
        done = 0
        while done < count:
            n = min(self.wpl, count - done)
            t0 = time.perf_counter()
            env = b.BatchedManEnv(self._dm[shape], n, device=self.device, dtype=torch.float32, seed=self.shape_seed(shape),
                                  sim_start=self.sim_start, sim_step=self.sim_step, world_offset=first + done, contact_mode=self.mode)
            if shape in self.tdamp:
                env.set_params(tendon_damping=torch.full((n,), float(self.tdamp[shape]), dtype=torch.float64, device=self.device))
            traj, k, st, touch = env.rollout(return_touch=True)
            if self.mask:
                env.mask_contact(traj, touch)
            torch.cuda.synchronize(self.device)
            self.seconds["simulate"] += time.perf_counter() - t0
            self.world_steps += n * (self.sim_start + self.sim_step * int(traj.shape[1]))
            yield traj, k, st
            done += n
            del env


DROP_BITS = 1 | 2 | 8      # diverged-and-reset, contact / candidate capacity exceeded, unsupported overlapping pair


def flag_counts(status):
    st = np.asarray(status)
    return {"diverged": int(((st & 1) != 0).sum()), "capacity": int(((st & 2) != 0).sum()), "unsupported": int(((st & 8) != 0).sum())}


def write_file(out_dir, stem, chunks, dataset, stats_fn=None, npz=False, drop_diverged=True):
    """chunks: iterable of (traj, stiffness, status) (numpy or torch).  A world whose state was reset by the NaN / 1e10
    check mid-episode (status bit 1; the reference's `except MujocoException: self.reset()`, ref: manenv.py:50-51) yields
    a trace with a discontinuity; a world that ran out of contact slots (bit 2) or met a pair outside the restated
    narrowphase set (bit 8) was simulated with contacts missing.  Such samples are dropped unless drop_diverged is False;
    either way they are counted per file.  Returns a summary dict."""
    trajs, ks, dev_chunks, ndiv = [], [], [], 0
    flags = {"diverged": 0, "capacity": 0, "unsupported": 0}
    for traj, k, st in chunks:
        st_np = st.cpu().numpy() if hasattr(st, "cpu") else np.asarray(st)
        bad = (st_np & DROP_BITS) != 0
        ndiv += int(bad.sum())
        for name, c in flag_counts(st_np).items():
            flags[name] += c
        if drop_diverged and bad.any():
            keep = np.nonzero(~bad)[0]
            if hasattr(traj, "cpu"):
                import torch
                keep = torch.from_numpy(keep).to(traj.device)
            traj, k = traj[keep], k[keep]
        if hasattr(traj, "cpu"):
            dev_chunks.append(traj)
            trajs.append(traj.double().cpu().numpy())
            ks.append(k.double().cpu().numpy())
        else:
            trajs.append(np.asarray(traj, dtype=np.float64))
            ks.append(np.asarray(k, dtype=np.float64))
    traj = np.concatenate(trajs, 0) if trajs else np.zeros((0, 200, 12))
    k = np.concatenate(ks, 0) if ks else np.zeros((0,))
    path = os.path.join(out_dir, stem + ".pickle")
    dataset.write_pickle(path, traj, k)
    extra = {}
    if stats_fn is not None and traj.shape[0] > 0:
        if dev_chunks:
            import torch
            mean, std = stats_fn(torch.cat(dev_chunks, 0).contiguous())     # on the device, before the host copy is used
        else:
            mean, std = stats_fn(traj)
        extra = {"mean": np.asarray(mean.cpu() if hasattr(mean, "cpu") else mean).reshape(-1),
                 "std": np.asarray(std.cpu() if hasattr(std, "cpu") else std).reshape(-1)}
    if traj.shape[0] > 0:
        edges, bins = dataset.feature_stats(traj, k)
        extra["bin_edges"] = edges
        for i, b in enumerate(bins):
            if b is not None:
                extra.update({"bin%d_n" % i: b["n"], "bin%d_mean" % i: b["mean"], "bin%d_std" % i: b["std"], "bin%d_peak" % i: b["peak"]})
    np.savez(os.path.join(out_dir, stem + ".stats.npz"), n=traj.shape[0], diverged=flags["diverged"], capacity=flags["capacity"],
             unsupported=flags["unsupported"], dropped=ndiv if drop_diverged else 0, **extra)
    if npz:
        dataset.write_npz(os.path.join(out_dir, stem + ".npz"), traj.astype(np.float32), k)
    return {"file": path, "samples": int(traj.shape[0]), "diverged": flags["diverged"], "capacity": flags["capacity"],
            "unsupported": flags["unsupported"], "dropped": ndiv if drop_diverged else 0}


def shard_parts(parts, rank, world):
    """Rank `rank` of `world` takes a contiguous slice of every part's world-id range (the sharding of SURVEY section 8e:
    no collective on the step path; per-world draws depend on the global id only, so the union over ranks, in rank order,
    is exactly the single-GPU result)."""
    out = []
    for shape, first, count in parts:
        lo, hi = (count * rank) // world, (count * (rank + 1)) // world
        if hi > lo:
            out.append((shape, first + lo, hi - lo))
    return out


def merge_shards(out_dir, stems, world, dataset, npz=False, remove=True):
    """The "final gather of dataset shards": concatenate `<stem>.rank<r>.pickle` of all ranks, part by part in rank order
    (every rank wrote its slice of each part as its own list of samples), into `<stem>.pickle`; the per-file statistics
    (`<stem>.stats.npz`: the reference's own numpy expression, ref: functions/utils.py:39-40, the per-stiffness-bin
    features and the summed status-flag counts of the ranks) are computed here, on the merged samples."""
    import pickle
    out = []
    for stem in stems:
        shards = []
        for r in range(world):
            with open(os.path.join(out_dir, "%s.rank%d.pickle" % (stem, r)), "rb") as f:
                shards.append(pickle.load(f))
        nparts = max(len(sh["part_sizes"]) for sh in shards)
        data, ks = [], []
        flags = {name: sum(int(sh.get("flags", {}).get(name, 0)) for sh in shards) for name in ("diverged", "capacity", "unsupported", "dropped")}
        for p in range(nparts):
            for sh in shards:
                a = sum(sh["part_sizes"][:p])
                b = a + (sh["part_sizes"][p] if p < len(sh["part_sizes"]) else 0)
                data += sh["data"][a:b]
                ks += sh["stiffness"][a:b]
        traj = np.stack(data) if data else np.zeros((0, 200, 12))
        dataset.write_pickle(os.path.join(out_dir, stem + ".pickle"), traj, ks)
        extra = {}
        if traj.shape[0] > 0:
            extra = {"mean": np.mean(traj, axis=(0, 1)), "std": np.std(traj, axis=(0, 1))}
            edges, bins = dataset.feature_stats(traj, np.asarray(ks))
            extra["bin_edges"] = edges
            for i, b in enumerate(bins):
                if b is not None:
                    extra.update({"bin%d_n" % i: b["n"], "bin%d_mean" % i: b["mean"], "bin%d_std" % i: b["std"], "bin%d_peak" % i: b["peak"]})
        np.savez(os.path.join(out_dir, stem + ".stats.npz"), n=traj.shape[0], **flags, **extra)
        if npz:
            dataset.write_npz(os.path.join(out_dir, stem + ".npz"), traj.astype(np.float32), np.asarray(ks))
        if remove:
            for r in range(world):
                os.remove(os.path.join(out_dir, "%s.rank%d.pickle" % (stem, r)))
        out.append(dict({"file": os.path.join(out_dir, stem + ".pickle"), "samples": len(data)}, **flags))
    return out


def regenerate_shard(out_dir, n_train, n_val, n_test, rollouts, rank, world, shapes=SHAPES, drop_diverged=True, noise_seed=None):
    """What one rank of a multi-GPU regeneration does: its slice of every part of every file -> `<stem>.rank<r>.pickle`
    (reference layout plus the per-part sample counts and status-flag counts the merge needs).  noise_seed: as in
    `regenerate` (the draw of a sample depends on its global world id only, so the merged file equals the single-GPU one).
    Returns the stems."""
    import pickle
    dataset = importlib.import_module(_PKG + ".dataset")
    stems = []
    for stem, parts in plan_files(n_train, n_val, n_test, shapes):
        kw = {"noise_seed": noise_seed} if (noise_seed is not None and stem.endswith("/train")) else {}
        data, ks, sizes = [], [], []
        flags = {"diverged": 0, "capacity": 0, "unsupported": 0, "dropped": 0}
        for shape, first, count in parts:
            n0 = len(data)
            for myshape, myfirst, mycount in shard_parts([(shape, first, count)], rank, world):
                for traj, k, st in rollouts(myshape, myfirst, mycount, **kw):
                    traj = traj.double().cpu().numpy() if hasattr(traj, "cpu") else np.asarray(traj, dtype=np.float64)
                    k = k.double().cpu().numpy() if hasattr(k, "cpu") else np.asarray(k, dtype=np.float64)
                    st = st.cpu().numpy() if hasattr(st, "cpu") else np.asarray(st)
                    for name, c in flag_counts(st).items():
                        flags[name] += c
                    keep = (st & DROP_BITS) == 0 if drop_diverged else np.ones(len(k), dtype=bool)
                    flags["dropped"] += int((~keep).sum())
                    d = dataset.to_reference_dict(traj[keep], k[keep])
                    data += d["data"]
                    ks += d["stiffness"]
            sizes.append(len(data) - n0)
        path = os.path.join(out_dir, "%s.rank%d.pickle" % (stem, rank))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump({"data": data, "stiffness": ks, "part_sizes": sizes, "flags": flags}, f)
        stems.append(stem)
    return stems


def regenerate_distributed(out_dir, n_train, n_val, n_test, rollouts, shapes=SHAPES, npz=False, drop_diverged=True, backend="nccl",
                           noise_seed=None, log=print):
    """Under torchrun (RANK / WORLD_SIZE / MASTER_* in the environment): one rank per GPU, world shards with no collective
    on the step path; the only exchange is the barrier before rank 0 concatenates the shard files.  Returns the merged
    file summaries on rank 0, None elsewhere."""
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    own = not dist.is_initialized()
    if own:
        if backend == "nccl":                              # one rank per GPU: the barrier's communicator lives on this rank's device
            import torch
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    dist.barrier()
    t0 = time.perf_counter()
    stems = regenerate_shard(out_dir, n_train, n_val, n_test, rollouts, rank, world, shapes=shapes, drop_diverged=drop_diverged,
                             noise_seed=noise_seed)
    dist.barrier()
    t1 = time.perf_counter()
    out = merge_shards(out_dir, stems, world, importlib.import_module(_PKG + ".dataset"), npz=npz) if rank == 0 else None
    dist.barrier()
    if rank == 0:
        for s_ in out:
            log("{file}: {samples} samples ({diverged} diverged, {capacity} capacity, {unsupported} unsupported; {dropped} dropped)".format(**s_))
        log("dataset tree written in %.1f s on %d ranks (shards %.1f s, merge on rank 0 %.1f s)" % (time.perf_counter() - t0, world, t1 - t0, time.perf_counter() - t1))
    if own:
        dist.destroy_process_group()
    return out


def regenerate(out_dir, n_train, n_val, n_test, rollouts, shapes=SHAPES, stats_fn=None, npz=False, drop_diverged=True, log=print,
               noise_seed=None):
    """rollouts(shape, first_world, count) -> iterable of (traj, stiffness, status) chunks.  Returns the per-file summaries.
    noise_seed: bake the trainer's noise augmentation into the ``*/train`` files (the reference adds it to training batches
    only, ref: functions/optimization.py:32-33); passed on as ``rollouts(..., noise_seed=...)``."""
    dataset = importlib.import_module(_PKG + ".dataset")
    out = []
    t0 = time.perf_counter()
    files = plan_files(n_train, n_val, n_test, shapes)
    if hasattr(rollouts, "prefetch"):
        rollouts.prefetch(files)            # one launch per shape instead of one per file part
    for stem, parts in files:
        kw = {"noise_seed": noise_seed} if (noise_seed is not None and stem.endswith("/train")) else {}

        def chunks():
            for shape, first, count in parts:
                yield from rollouts(shape, first, count, **kw)
        t_file = time.perf_counter()
        s = write_file(out_dir, stem, chunks(), dataset, stats_fn=stats_fn, npz=npz, drop_diverged=drop_diverged)
        s["parts"] = parts
        out.append(s)
        s["seconds"] = time.perf_counter() - t_file
        log("{file}: {samples} samples ({diverged} diverged, {capacity} capacity, {unsupported} unsupported; {dropped} dropped) {seconds:.2f} s".format(**s))
    total = time.perf_counter() - t0
    log("dataset tree written in %.1f s" % total)
    sec = getattr(rollouts, "seconds", None)
    if sec is not None:
        import json
        log(json.dumps({"regenerate": {"seconds_total": round(total, 3), "seconds_models": round(sec["model"], 3),
                                       "seconds_rollouts_incl_launch_setup": round(sec["simulate"], 3),
                                       "seconds_copy_pickle_stats": round(total - sec["model"] - sec["simulate"], 3),
                                       "samples": int(sum(f["samples"] for f in out)), "world_steps": int(rollouts.world_steps),
                                       "world_steps_per_s_end_to_end": rollouts.world_steps / total if total > 0 else None}}))
    return out


def build_parser():
    p = ArgumentParser(description=__doc__.split("\n")[0])
    p.add_argument("--out", type=str, default="./data/experiments")
    p.add_argument("--train", type=int, default=4096)
    p.add_argument("--val", type=int, default=512)
    p.add_argument("--test", type=int, default=512)
    for s in SHAPES:
        p.add_argument("--" + s, type=str, default=None, help="MJCF (.xml) or compiled blob (.sgm) of the %s scene" % s)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--worlds-per-launch", type=int, default=16384)
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--mask-contact", action="store_true")
    p.add_argument("--contact-mode", choices=["intended", "reference"], default="intended")
    p.add_argument("--keep-diverged", action="store_true")
    p.add_argument("--npz", action="store_true")
    p.add_argument("--noise-seed", type=int, default=None,
                   help="bake noised_modality (sigma 0.7 / 0.06) into the */train files on the device")
    p.add_argument("--tendon-damping", type=str, nargs="*", default=[], metavar="SHAPE=VALUE",
                   help="volume-tendon damper per shape, e.g. softball=50 softcylinder=50 (committed value: 100, at which these "
                        "two models run away under the restated semantics and every sample is dropped; DESIGN.md section 4)")
    p.add_argument("--sim-step", type=int, default=7)
    p.add_argument("--sim-start", type=int, default=1)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    paths = {s: getattr(args, s) for s in SHAPES if getattr(args, s)}
    if not paths:
        raise SystemExit("give at least one of --softball / --softbox / --softcylinder")
    fn = importlib.import_module(_PKG + ".functions")
    tdamp = {}
    for item in args.tendon_damping:
        shape, _, value = item.partition("=")
        if shape not in SHAPES or not value:
            raise SystemExit("--tendon-damping takes SHAPE=VALUE with SHAPE in %s" % (SHAPES,))
        tdamp[shape] = float(value)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        roll = DeviceRollouts(paths, seed=args.seed, worlds_per_launch=args.worlds_per_launch, device="cuda:%d" % local,
                              mask_contact=args.mask_contact, contact_mode=args.contact_mode, sim_start=args.sim_start, sim_step=args.sim_step,
                              tendon_damping=tdamp)
        return regenerate_distributed(args.out, args.train, args.val, args.test, roll, shapes=tuple(s for s in SHAPES if s in paths),
                                      npz=args.npz, drop_diverged=not args.keep_diverged, backend="nccl", noise_seed=args.noise_seed)
    roll = DeviceRollouts(paths, seed=args.seed, worlds_per_launch=args.worlds_per_launch, device=args.device,
                          mask_contact=args.mask_contact, contact_mode=args.contact_mode, sim_start=args.sim_start, sim_step=args.sim_step,
                          tendon_damping=tdamp)
    return regenerate(args.out, args.train, args.val, args.test, roll, shapes=tuple(s for s in SHAPES if s in paths),
                      stats_fn=fn.channel_mean_std, npz=args.npz, drop_diverged=not args.keep_diverged, noise_seed=args.noise_seed)


if __name__ == "__main__":
    main()
