#!/usr/bin/env python
"""The headline C2 frame (1920x1080, 64 spp) through the IN-LIBRARY multi-GPU path (ne_b200_create_multi, csrc/ne_multi.cu):
ONE process, N devices, scene replicas, sample-index partition, one fused peer-memory reduce + resolve kernel on the first
device. Prints one JSON line per N (strong scaling: the same frame at every N). Timing: host wall clock around
ne_b200_multi_render + ne_b200_multi_resolve (the resolve joins every device and copies the frame to pinned host memory),
every device idle before.   usage: python tools/run_inproc_multi.py [N ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402  (pinned host memory + device synchronisation only)
import bench  # noqa: E402
from narvalengine_b200.engine import MultiContext  # noqa: E402


def main():
    ns = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
    ndev = torch.cuda.device_count()
    b, cam_params, grid, _ = bench.build_scene(pin=True)
    tm = bench.pinned_like(np.empty((bench.H, bench.W, 3), np.float32))
    base = None
    for n in ns:
        if n > ndev:
            continue
        m = MultiContext(list(range(n)))
        m.upload(b)
        cam = cam_params.make(bench.W / bench.H, m.lib)
        for i in range(3):
            m.render_frame(cam, bench.W, bench.H, bench.SPP, bench.BOUNCES, seed=1 + i, tonemapped=tm)
        ts, ts_render = [], []
        for i in range(5):
            for d in range(n):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            m.render(cam, bench.W, bench.H, bench.SPP, bench.BOUNCES, seed=10 + i)
            t1 = time.perf_counter()
            m.resolve(tm, None)
            ts.append(time.perf_counter() - t0)
            ts_render.append(t1 - t0)
        t = float(np.mean(ts))
        v = bench.W * bench.H * bench.SPP / t / 1e6
        base = base or v
        print(json.dumps({"path": "in-library multi-GPU (one process)", "n_gpus": n, "Mpaths_per_s": v, "ms_per_frame": t * 1e3, "scaling": "strong",
                          "efficiency_vs_first": v / (base * n / ns[0]), "host_ms_to_enqueue_all_devices": float(np.mean(ts_render)) * 1e3,
                          "peer_access": [m.peer_access(r) for r in range(n)], "includes": "render on N devices + fused reduce/resolve + D2H of the tone-mapped frame",
                          "mean": float(tm.mean())}), flush=True)
        m.close()


if __name__ == "__main__":
    main()
