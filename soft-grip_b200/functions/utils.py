"""The dataset statistics of ``create_tf_generators`` (ref: functions/utils.py:39-40):
``train_mean = np.mean(train_x, axis=(0, 1), keepdims=True)``, ``train_std = np.std(train_x, axis=(0, 1), keepdims=True)``
computed on the device from the trajectory tensor a rollout produced, without a host round trip."""
from .._lib import check, lib
from ._traj import ptr, traj_args


def channel_mean_std(train_x, workspace=None, out=None):
    """train_x: CUDA tensor (N, T, C) -> (mean, std), float64 tensors of shape (1, 1, C) (population std, like np.std).
    workspace / out: optional preallocated float64 CUDA tensors (``stats_workspace(train_x)`` / shape (2, C)) for callers that
    run the statistics repeatedly and do not want two allocations per call."""
    torch, nrows, nchan, prec, dev, stream = traj_args(train_x)
    L = lib()
    if workspace is None:
        workspace = stats_workspace(train_x)
    nbytes = workspace.numel() * 8
    if out is None:
        out = torch.empty((2, nchan), dtype=torch.float64, device=train_x.device)
    check(L.sg_traj_channel_stats(ptr(train_x), nrows, nchan, prec, dev, ptr(out[0]), ptr(out[1]), ptr(workspace), nbytes, stream))
    shape = (1,) * (train_x.dim() - 1) + (nchan,)
    return out[0].reshape(shape), out[1].reshape(shape)


def stats_workspace(train_x):
    """The scratch tensor ``channel_mean_std`` needs for a tensor of this shape (sg_traj_stats_workspace_bytes)."""
    torch, nrows, nchan, prec, dev, stream = traj_args(train_x)
    nbytes = check(lib().sg_traj_stats_workspace_bytes(nrows, nchan, dev))
    return torch.empty(max(1, nbytes // 8), dtype=torch.float64, device=train_x.device)
