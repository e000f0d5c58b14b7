#!/usr/bin/env python
"""BASELINE configs[4]: the mixed scene (500k-triangle mesh + rectangles + sphere light + sparse cloud) at 3840x2160,
1024 spp, split by SAMPLE INDEX across the GPUs of one box, per-GPU accumulation buffers summed onto rank 0 by one NCCL
reduce. Launch like bench.py:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/run_c5_multi.py

Prints one JSON line on rank 0 (device-timed, max over ranks). --spp / --width / --height shrink it for a smoke run."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import scenes
    from imgmetrics import luminance
    from narvalengine_b200.engine import Context
    from narvalengine_b200.multigpu import PartitionedFrame, alias_accum, sample_range

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, spp, bounces = args.width, args.height, args.spp, 6
    t0 = time.time()
    b = scenes.mixed_scene(res=(256, 176, 304), mesh_n=500)
    t_build = time.time() - t0
    ctx = Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.upload(b)
    ctx.set_camera(scenes.MIXED_CAMERA.make(W / H, ctx.lib))
    ctx.render(W, H, 0, 0, bounces)
    frame = PartitionedFrame(ctx, alias_accum(ctx, local), rank, world, dist if world > 1 else None)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    frame.render(W, H, max(world, spp // 8), bounces, seed=1)  # warm-up (allocates the wavefront pool)
    sync()
    ctx.counters_reset()
    ms = 0.0
    for i in range(args.steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        frame.render(W, H, spp, bounces, seed=2 + i)
        e1.record(stream)
        sync()
        ms += e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    c = ctx.counters()
    rays = torch.tensor([float(c.extend_rays + c.shadow_rays)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rays)
    if rank == 0:
        lin = np.zeros((H, W, 3), np.float32)
        ctx.read_linear(W, H, lin)
        print(json.dumps({"config": "c5 (BASELINE configs[4] shape: 500k-triangle mesh + analytic primitives + sparse cloud 256x176x304)",
                          "resolution": [W, H], "spp": spp, "n_gpus": world, "partition": "sample index, NCCL reduce of the accumulation buffers onto rank 0",
                          "samples_rank0": list(sample_range(0, world, spp)), "Mpaths_per_s": W * H * spp / ms / 1e3,
                          "Mrays_per_s": float(rays.item()) / args.steps / ms / 1e3, "frame_ms": ms, "scene_build_s": t_build,
                          "mean_luminance": float(luminance(lin).mean()), "finite": bool(np.isfinite(lin).all())}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
