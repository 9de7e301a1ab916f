"""soft-grip_b200: B200-native batched simulator behind the reference's ManEnv API.

The package name contains a hyphen (it mirrors the reference repository's name), so import it with
``importlib.import_module("soft-grip_b200")`` or put this directory on ``sys.path`` and use the
drop-in ``environment`` package inside it (``from environment import ManEnv``).
"""
from . import mjcf  # noqa: F401
from ._lib import SoftGripError, LIB_PATH  # noqa: F401
