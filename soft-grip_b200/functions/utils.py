"""The dataset statistics of ``create_tf_generators`` (ref: functions/utils.py:39-40):
``train_mean = np.mean(train_x, axis=(0, 1), keepdims=True)``, ``train_std = np.std(train_x, axis=(0, 1), keepdims=True)``
computed on the device from the trajectory tensor a rollout produced, without a host round trip."""
from .._lib import check, lib
from ._traj import ptr, traj_args


def channel_mean_std(train_x):
    """train_x: CUDA tensor (N, T, C) -> (mean, std), float64 tensors of shape (1, 1, C) (population std, like np.std)."""
    torch, nrows, nchan, prec, dev, stream = traj_args(train_x)
    L = lib()
    nbytes = check(L.sg_traj_stats_workspace_bytes(nrows, nchan, dev))
    ws = torch.empty(max(1, nbytes // 8), dtype=torch.float64, device=train_x.device)
    out = torch.empty((2, nchan), dtype=torch.float64, device=train_x.device)
    check(L.sg_traj_channel_stats(ptr(train_x), nrows, nchan, prec, dev, ptr(out[0]), ptr(out[1]), ptr(ws), nbytes, stream))
    shape = (1,) * (train_x.dim() - 1) + (nchan,)
    return out[0].reshape(shape), out[1].reshape(shape)
