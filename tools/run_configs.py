#!/usr/bin/env python
"""BASELINE.json's configurations at (or near) full size through the GPU backend, with a small-frame check against the
reference oracle (oracle/_ref) where the CPU finishes in seconds. Prints one JSON line per configuration.

  python tools/run_configs.py [c1] [c2] [c3] [c4] [c5] [--check] [--spp N]

c1  Cornell box (rectangles + sphere emitter, GGX, area light)      512x512x16       
c2  heterogeneous 256^3 volume + point light                         1920x1080x64     (C2, the bench.py workload)
c3  sparse procedural cloud handed over as OpenVDB-style 8^3 LEAVES  1920x1080x64     (C3 shape; grid 384x264x456, "sun" =
    big distant rectangle emitter)
c4  2M-triangle displaced grid mesh, GGX, two area lights             1920x1080x64     (C4)
c5  mixed: 500k-triangle mesh + rectangles + sphere light + cloud     1920x1080x64     (C5 shape at 1080p on one GPU)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import scenes  # noqa: E402
from imgmetrics import luminance, rel_mse  # noqa: E402
from narvalengine_b200.engine import Context  # noqa: E402


def c1():
    return scenes.cornell_c1(), scenes.CORNELL_CAMERA, 512, 512, 16


def c2():
    return scenes.c2_scene(), scenes.C2_CAMERA, 1920, 1080, 64


def c3():
    res = (384, 264, 456)
    grid = scenes.cloud_density(res, seed=7)
    o, v = scenes.dense_to_leaves(grid)
    b = scenes.SceneBuilder()
    vol = b.add_volume_leaves(res, o, v)
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 60.0, vol, "hg", 0.0)
    b.add_emitter("sun", (900, 850, 700))
    b.add_volume("cloud", (0, 0, 0), (0, 0, 0), (15.9, 9.51, 13.5))
    b.add_rectangle("sun", (20, 40, -10), (-60, 25, 0), (30, 30, 1))
    return b, scenes.CameraParams((0, 2, -30), (0, 0, 0), 40.0), 1920, 1080, 64


def c4(n=1000):
    b = scenes.mesh_scene(n=n)
    return b, scenes.MESH_CAMERA, 1920, 1080, 64


def c5():
    b = scenes.mixed_scene(res=(256, 176, 304), mesh_n=500)
    return b, scenes.MIXED_CAMERA, 1920, 1080, 64


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    check = "--check" in sys.argv
    spp_override = int(sys.argv[sys.argv.index("--spp") + 1]) if "--spp" in sys.argv else None
    ctx = Context(0)
    for name in args or ["c1", "c2", "c3", "c4", "c5"]:
        t0 = time.time()
        b, cam, W, H, spp = globals()[name]()
        spp = spp_override or spp
        t_build = time.time() - t0
        t0 = time.time()
        ctx.upload(b)
        t_up = time.time() - t0
        camera = cam.make(W / H, ctx.lib)
        lin = np.zeros((H, W, 3), np.float32)
        ctx.render_frame(camera, W, H, spp, 6, 1, 0, None, lin)  # warm-up at full size: the wavefront pool is sized by the work
        ctx.counters_reset()
        t0 = time.time()
        ctx.render_frame(camera, W, H, spp, 6, 2, 0, None, lin)
        dt = time.time() - t0
        c = ctx.counters()
        line = {"config": name, "resolution": [W, H], "spp": spp, "Mpaths_per_s": W * H * spp / dt / 1e6,
                "Mrays_per_s": (c.extend_rays + c.shadow_rays) / dt / 1e6, "frame_ms": dt * 1e3, "device_ms": c.ms_render,
                "upload_s": t_up, "scene_build_s": t_build, "mean_luminance": float(luminance(lin).mean()), "finite": bool(np.isfinite(lin).all()),
                "kernel_ms": {"volume": c.ms_volume_kernel, "extend_shadow": c.ms_extend_kernel, "shade": c.ms_shade_kernel, "generate_plan": c.ms_other_kernel},
                "counters": {k: int(getattr(c, k)) for k in ("paths", "extend_rays", "shadow_rays", "delta_steps", "ratio_steps", "brick_visits",
                                                              "bvh_nodes", "tri_tests", "scatter_events", "surface_events", "wavefront_iterations")}}
        if check:
            from refclient import RefOracle
            w, h, s = 96, 54 if W != H else 96, 256
            a = np.zeros((h, w, 3), np.float32)
            ctx.render_frame(cam.make(w / h, ctx.lib), w, h, s, 6, 3, 0, None, a)
            t0 = time.time()
            sc = RefOracle().scene(b)
            ref, secs = sc.render(cam, w, h, s, 6, seed=5, threads=os.cpu_count() or 1)
            line["oracle_check"] = {"resolution": [w, h], "spp": s, "rel_mse": float(rel_mse(a, ref)),
                                    "luminance_ratio": float(luminance(a).mean() / luminance(ref).mean()), "oracle_s": time.time() - t0}
            sc.close()
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
