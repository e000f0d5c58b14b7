"""Host-side breakdown of the end-to-end frame: ne_b200_scene_upload vs ne_b200_render_frame (headline C2 frame)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import bench
from narvalengine_b200.engine import Context
b, cam_params, grid, _ = bench.build_scene()
ctx = Context(0)
desc = b.desc()
W, H = 1920, 1080
cam = cam_params.make(W / H, ctx.lib)
tm = np.empty((H, W, 3), np.float32)
for i in range(int(os.environ.get("N", "5"))):
    t0 = time.perf_counter(); ctx.upload(desc); t1 = time.perf_counter()
    ctx.render_frame(cam, W, H, 64, 6, 100 + i, 0, tm, None); t2 = time.perf_counter()
    c = ctx.counters()
    print("upload %.2f ms (lib %.2f)  render_frame %.2f ms  (device render %.2f)" % ((t1 - t0) * 1e3, c.ms_upload, (t2 - t1) * 1e3, c.ms_render)); ctx.counters_reset()
