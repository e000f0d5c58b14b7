"""Quick check that the render graph (WHILE node + device-side loop condition) runs: one small frame of a volume scene and
of a mesh scene through the graph and through the host-driven loop, compared. `small`: tiny frames (compute-sanitizer)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import scenes  # noqa: E402
from narvalengine_b200.engine import Context  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "small"
ctx = Context(0)
cases = [("volume", scenes.noise_volume_scene(res=(48, 48, 48), density=40.0, light="point"), scenes.CameraParams((0, 1, -6), (0, 1, 0), 45.0)),
         ("mixed", scenes.mixed_scene(), scenes.MIXED_CAMERA), ("cornell", scenes.cornell_c1(), scenes.CORNELL_CAMERA)]
W, H, spp = (32, 24, 4) if small else (192, 128, 16)
for name, b, cp in cases:
    ctx.upload(b)
    cam = cp.make(W / H, ctx.lib)
    out = {}
    for mode in ("graph", "host"):
        if mode == "host":
            os.environ["NE_B200_HOST_LOOP"] = "1"
        else:
            os.environ.pop("NE_B200_HOST_LOOP", None)
        ctx.counters_reset()
        lin = np.zeros((H, W, 3), np.float32)
        t = time.time()
        ctx.render_frame(cam, W, H, spp, 6, 3, 0, None, lin)
        dt = time.time() - t
        c = ctx.counters()
        out[mode] = lin
        print(f"{name} {mode}: mean {lin.mean():.5f} wall {dt * 1e3:.1f} ms iters {c.wavefront_iterations} launches {c.kernel_launches} "
              f"ms render {c.ms_render:.3f} vol {c.ms_volume_kernel:.3f} trace {c.ms_extend_kernel:.3f} shade {c.ms_shade_kernel:.3f} other {c.ms_other_kernel:.3f}",
              flush=True)
    d = np.abs(out["graph"] - out["host"]).max() / max(1e-9, out["host"].mean())
    print(f"{name}: max |graph - host| / mean = {d:.2e}")
    assert d < 1e-2 or small
ctx.close()
print("sanity ok")
