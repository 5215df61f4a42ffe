"""sat_b200: B200-native (sm_100a) implementation of the Show-Attend-and-Tell decoder hot path,
behind the reference's model.py API.  See DESIGN.md / INTEGRATION.md."""
from . import _lib  # noqa: F401
from ._lib import SatError, build, launch_count  # noqa: F401
