"""anim-nerf_b200: B200-native (sm_100a) implementation of the Anim-NeRF per-ray
rendering hot path behind the reference's Python API.  See DESIGN.md."""
from . import synthetic  # noqa: F401
from . import body_model  # noqa: F401
