"""Dataset generation driver with the reference's command line (ref: create_dataset.py:83-93).

``python create_dataset.py --mujoco-model-paths A.xml [B.xml ...]`` squeezes every listed object once
per ``NUM_EPISODES`` with the B200 ``ManEnv`` and pickles ``{"data": [...], "stiffness": [...]}`` exactly
like the reference driver (protocol: ref: create_dataset.py:33-72, file layout: ref: create_dataset.py:75-78).

The episode protocol lives in :func:`episode_rows`; :func:`log_into_file` keeps the reference's entry
point name and argument object.  ``--batched N`` (an extension) produces N episodes per model from one
rollout-kernel launch per model instead of driving single environments from Python.
"""
import importlib
import os
import sys
from argparse import ArgumentParser

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)
from environment import ManEnv  # noqa: E402  (drop-in package next to this file)

_PKG = os.path.basename(_HERE)

# episode constants, same names and values as the reference module globals (ref: create_dataset.py:14-17)
NUM_EPISODES = 1
MAX_ITER_PER_EP = 160
OPEN_CLOSE_DIV = 80
START_STEP = 40


def episode_rows(env, mask_contact):
    """Yield the sensor rows of one squeeze episode driven through the ManEnv verbs.

    START_STEP settle steps with the ctrl left at zero by reset(), close_hand(), then MAX_ITER_PER_EP
    steps with a toggle_grip() whenever the step index is a positive multiple of OPEN_CLOSE_DIV.
    """
    def one_row():
        readings, contact = env.step()
        if mask_contact and not contact:
            readings = np.zeros_like(readings)
        return readings

    for _ in range(START_STEP):
        yield one_row()
    env.close_hand()
    for i in range(MAX_ITER_PER_EP):
        env.render()
        if i > 0 and i % OPEN_CLOSE_DIV == 0:
            env.toggle_grip()
        yield one_row()


def log_into_file(args):
    """Reference entry point (ref: create_dataset.py:20): one ManEnv, NUM_EPISODES episodes per model path,
    next model loaded with ``load_env`` after each group, one pickle at the end."""
    paths = args.mujoco_model_paths
    assert type(paths) is list
    dataset = importlib.import_module(_PKG + ".dataset")
    env = ManEnv(**ManEnv.get_std_spec(args))
    traces, labels = [], []
    which = 0
    for ep in range(NUM_EPISODES * len(paths)):
        label = env.reset()
        rows = [r for r in episode_rows(env, args.mask_contact) if r is not None]
        traces.append(np.array(rows))            # array(), not asarray(): the rows must be copied
        labels.append(label)
        if len(paths) > 1 and (ep + 1) % NUM_EPISODES == 0:
            which += 1
            if which > len(paths):               # same off-by-one as the reference
                which = 0
            env.load_env(which)
    os.makedirs(args.data_folder, exist_ok=True)
    out = os.path.join(args.data_folder, "{}.pickle".format(args.data_name))
    dataset.write_pickle(out, np.stack(traces), labels)
    print("Total number of samples: {0}".format(len(traces)))
    return out


def log_into_file_batched(args):
    """Extension: ``args.batched`` episodes per model, each model in a single rollout launch."""
    import torch
    batched = importlib.import_module(_PKG + ".batched")
    dataset = importlib.import_module(_PKG + ".dataset")
    trajs, ks = [], []
    for path in args.mujoco_model_paths:
        env = batched.BatchedManEnv(path, args.batched, dtype=torch.float32, seed=args.seed,
                                    sim_start=args.sim_start, sim_step=args.sim_step)
        sched = batched.default_schedule(env.nu, START_STEP, MAX_ITER_PER_EP, OPEN_CLOSE_DIV)
        traj, k, st, touch = env.rollout(schedule=sched, return_touch=True)
        if args.mask_contact:
            env.mask_contact(traj, touch)            # on the device, in place
        traj = traj.double().cpu().numpy()
        ndiv = int(((st.cpu().numpy() & batched.ST_DIVERGED) != 0).sum())
        if ndiv:
            print("warning: {} of {} worlds diverged and were reset mid-episode".format(ndiv, args.batched))
        trajs.append(traj)
        ks.append(k.cpu().numpy())
    out = os.path.join(args.data_folder, "{}.pickle".format(args.data_name))
    dataset.write_pickle(out, np.concatenate(trajs, 0), np.concatenate(ks, 0))
    print("Total number of samples: {0}".format(sum(t.shape[0] for t in trajs)))
    return out


def build_parser():
    parser = ArgumentParser()
    # reference flags (ref: create_dataset.py:84-92); type=bool keeps the reference's "any non-empty string
    # is True" behaviour
    parser.add_argument('--sim-step', type=int, default=7)
    parser.add_argument('--vis', type=bool, default=True)
    parser.add_argument('--mask-contact', type=bool, default=False)
    parser.add_argument('--sim-start', type=int, default=1)
    parser.add_argument('--data-folder', type=str, default="./data/dataset/testing_datasets")
    parser.add_argument('--data-name', type=str, default="dataset_all_shapes")
    parser.add_argument('--mujoco-model-paths', nargs="+", required=True)
    # extensions
    parser.add_argument('--batched', type=int, default=0, help="episodes per model through the rollout kernel")
    parser.add_argument('--seed', type=int, default=0)
    return parser


if __name__ == '__main__':
    cli_args, _ = build_parser().parse_known_args()
    (log_into_file_batched if cli_args.batched > 0 else log_into_file)(cli_args)
