#!/usr/bin/env python
"""BASELINE configs[2] at its full size: a WDAS-cloud-scale sparse volume (1000 x 700 x 1200 voxels) handed
over as OpenVDB-style 8^3 LEAVES (origin + 512 floats, inactive leaves omitted), big distant rectangle "sun", 1920x1080,
256 spp. The density (union of ~200 seeded spheres eroded by trilinear lattice fBm) is generated slab by slab with
torch on the GPU - scene SYNTHESIS only, the 3.4 GB dense grid never exists - and copied to host leaf arrays, which
then go through the product's normal path: ne_b200_scene_upload (host brick builder for leaf input) + render.

  python tools/run_c3_full.py [--scale 1.0] [--spp 256]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402


def cloud_leaves(W, H, D, seed=7, n_spheres=200, device="cuda"):
    """-> origins [n,3] int32 (x,y,z multiples of 8), values [n,8,8,8] float32 indexed [z,y,x], active voxel fraction."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cpu").manual_seed(seed)
    aspect = torch.tensor([W, H, D], dtype=torch.float32) / max(W, H, D)
    c = (torch.rand(n_spheres, 3, generator=g) - 0.5) * aspect * 0.8          # centres, in units of the longest axis
    c[:, 1] *= 0.6                                                            # flattened: a cumulus deck
    r = 0.04 + 0.06 * torch.rand(n_spheres, generator=g)
    lattices = [torch.rand(1, 1, f + 1, f + 1, f + 1, generator=g) for f in (6, 12, 24, 48)]
    c, r = c.to(device), r.to(device)
    lattices = [lt.to(device) for lt in lattices]
    Wp, Hp = (W + 7) // 8 * 8, (H + 7) // 8 * 8
    xs = (torch.arange(Wp, device=device, dtype=torch.float32) + 0.5) / max(W, H, D) - 0.5 * float(aspect[0])
    ys = (torch.arange(Hp, device=device, dtype=torch.float32) + 0.5) / max(W, H, D) - 0.5 * float(aspect[1])
    gx = ((torch.arange(Wp, device=device, dtype=torch.float32) + 0.5) / W * 2 - 1).clamp(-1, 1)
    gy = ((torch.arange(Hp, device=device, dtype=torch.float32) + 0.5) / H * 2 - 1).clamp(-1, 1)
    origins, values, active_vox = [], [], 0
    for z0 in range(0, D, 8):
        zi = torch.arange(z0, z0 + 8, device=device, dtype=torch.float32)
        zs = (zi + 0.5) / max(W, H, D) - 0.5 * float(aspect[2])
        gz = ((zi + 0.5) / D * 2 - 1).clamp(-1, 1)
        Z, Y, X = torch.meshgrid(zs, ys, xs, indexing="ij")
        dens = torch.zeros_like(X)
        for k in range(n_spheres):  # union of soft spheres
            d2 = (X - c[k, 0]) ** 2 + (Y - c[k, 1]) ** 2 + (Z - c[k, 2]) ** 2
            dens = torch.maximum(dens, (1 - d2 / (r[k] * r[k])).clamp_(0, 1))
        grid = torch.stack(torch.meshgrid(gz, gy, gx, indexing="ij")[::-1], -1)[None]  # (1,8,Hp,Wp,3) as (x,y,z)
        n, amp, tot = torch.zeros_like(X), 1.0, 0.0
        for lt in lattices:
            n += amp * F.grid_sample(lt, grid, mode="bilinear", padding_mode="border", align_corners=True)[0, 0]
            tot += amp
            amp *= 0.5
        dens = (dens * 1.6 - (n / tot) * 0.9).clamp_(0, 1)                    # fBm erosion
        dens[zi >= D] = 0
        dens[:, H:, :] = 0
        dens[:, :, W:] = 0
        blocks = dens.reshape(8, Hp // 8, 8, Wp // 8, 8).permute(1, 3, 0, 2, 4).reshape(-1, 512)
        act = blocks.amax(1) > 0
        idx = act.nonzero()[:, 0]
        if len(idx):
            by, bx = idx // (Wp // 8), idx % (Wp // 8)
            origins.append(torch.stack([bx * 8, by * 8, torch.full_like(bx, z0)], 1).to(torch.int32).cpu())
            values.append(blocks[idx].cpu())
            active_vox += int((blocks[idx] > 0).sum())
    o = torch.cat(origins).numpy()
    v = torch.cat(values).numpy().reshape(-1, 8, 8, 8)
    return np.ascontiguousarray(o), np.ascontiguousarray(v), active_vox / float(W * H * D)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of 1000x700x1200")
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    import scenes
    from imgmetrics import luminance
    from narvalengine_b200.engine import Context
    W3, H3, D3 = (max(8, int(round(x * args.scale))) for x in (1000, 700, 1200))
    t0 = time.time()
    o, v, frac = cloud_leaves(W3, H3, D3)
    t_gen = time.time() - t0
    b = scenes.SceneBuilder()
    vol = b.add_volume_leaves((W3, H3, D3), o, v)
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 60.0, vol, "hg", 0.0)
    b.add_emitter("sun", (900, 850, 700))
    b.add_volume("cloud", (0, 0, 0), (0, 0, 0), (15.9, 9.51, 13.5))   # cloudShowCase.json:41-50
    b.add_rectangle("sun", (20, 40, -10), (-60, 25, 0), (30, 30, 1))
    cam = scenes.CameraParams((0, 2, -30), (0, 0, 0), 40.0)
    W, H, spp = args.width, args.height, args.spp
    ctx = Context(0)
    t0 = time.time()
    ctx.upload(b)
    t_up = time.time() - t0
    camera = cam.make(W / H, ctx.lib)
    lin = np.zeros((H, W, 3), np.float32)
    ctx.render_frame(camera, W, H, spp, 6, 1, 0, None, lin)  # warm-up at full size
    ctx.counters_reset()
    t0 = time.time()
    ctx.render_frame(camera, W, H, spp, 6, 2, 0, None, lin)
    dt = time.time() - t0
    c = ctx.counters()
    print(json.dumps({"config": "c3 (BASELINE configs[2]: sparse cloud from OpenVDB-style 8^3 leaves, sun rectangle)", "grid": [W3, H3, D3],
                      "leaves": int(len(o)), "leaf_bytes": int(v.nbytes), "active_voxel_fraction": frac, "resolution": [W, H], "spp": spp,
                      "Mpaths_per_s": W * H * spp / dt / 1e6, "Mrays_per_s": (c.extend_rays + c.shadow_rays) / dt / 1e6, "frame_ms": dt * 1e3,
                      "device_ms": c.ms_render, "upload_s": t_up, "scene_synthesis_s": t_gen, "mean_luminance": float(luminance(lin).mean()),
                      "finite": bool(np.isfinite(lin).all()),
                      "kernel_ms": {"volume": c.ms_volume_kernel, "extend_shadow": c.ms_extend_kernel, "shade": c.ms_shade_kernel},
                      "counters": {k: int(getattr(c, k)) for k in ("paths", "extend_rays", "shadow_rays", "delta_steps", "ratio_steps", "brick_visits",
                                                                    "scatter_events", "wavefront_iterations")}}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
