"""Device-side counterparts of the two dataset operations of the reference's ``functions`` package that touch the
sensor trajectories (ref: functions/optimization.py:6-14, functions/utils.py:39-40).  The nets, losses and training
loops of that package are out of scope."""
from .optimization import noised_modality, SIGMA_ACC, SIGMA_GYRO  # noqa: F401
from .utils import channel_mean_std, stats_workspace  # noqa: F401
