"""``noised_modality`` of the reference (ref: functions/optimization.py:6-14) as one pass over the trajectory tensor in
HBM: accelerometer channels ``[..., :6] += N(0, 0.7)``, gyro channels ``[..., 6:] += N(0, 0.06)``.

The reference draws from TensorFlow's global, unseeded generator; here the draw of an element is a pure function of
``(seed, flat element index)`` (Philox4x32-10 + Box-Muller, see csrc/sg_traj.cuh), so augmented batches are reproducible
and independent of how the dataset is sharded.  Optionally fused with the ``(x - mean) / std`` that the reference
applies right after (ref: functions/optimization.py:38)."""
from .._lib import check, lib
from ._traj import ptr, traj_args

SIGMA_ACC, SIGMA_GYRO = 0.7, 0.06       # ref: functions/optimization.py:9,12


def noised_modality(data, seed=0, sigma_acc=SIGMA_ACC, sigma_gyro=SIGMA_GYRO, nacc=None, mean=None, std=None, out=None, first_row=0):
    """data: CUDA tensor (N, T, C) (C = 12: 6 accelerometer + 6 gyro channels) -> noised copy (or ``out``, which may be
    ``data`` itself for the reference's in-place ``+=``).  mean/std: optional (.., C) tensors for the fused
    standardisation.  first_row: row index (N*T flattened) of ``data[0, 0]`` inside the whole dataset tensor when ``data``
    is a shard of it -- shards noised separately then equal the rows of the whole tensor noised in one call."""
    torch, nrows, nchan, prec, dev, stream = traj_args(data)
    if nacc is None:
        nacc = nchan // 2               # ref: functions/optimization.py:8 splits 12 channels at 6
    if out is None:
        out = torch.empty_like(data)
    elif out.shape != data.shape or out.dtype != data.dtype or out.device != data.device or not out.is_contiguous():
        raise ValueError("out must match data")
    if (mean is None) != (std is None):
        raise ValueError("mean and std must be given together")
    m = s = None
    if mean is not None:
        m = torch.as_tensor(mean, dtype=torch.float64, device=data.device).reshape(-1).contiguous()
        s = torch.as_tensor(std, dtype=torch.float64, device=data.device).reshape(-1).contiguous()
        if m.numel() != nchan or s.numel() != nchan:
            raise ValueError("mean/std must have one entry per channel")
    check(lib().sg_traj_add_noise(ptr(data), ptr(out), nrows, int(first_row), nchan, int(nacc), float(sigma_acc), float(sigma_gyro),
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(m), ptr(s), prec, dev, stream))
    return out
