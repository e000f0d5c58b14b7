"""narvalengine_b200 — B200-native (sm_100a CUDA) path-tracing backend behind NarvalEngine's
OfflineEngine / Integrator seam. See DESIGN.md and include/ne_b200.h."""
from . import abi  # noqa: F401
from .scene import SceneBuilder, CameraParams  # noqa: F401
