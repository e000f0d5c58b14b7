#!/usr/bin/env python
"""bench.py — Mpaths/s of the path-tracing hot path on BASELINE.json's headline configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): one heterogeneous 256^3 density volume generated with the
reference tree's vendored FastNoise exactly as SURVEY 8d specifies (SimplexFractal, seed 1337, frequency 4/256, 5 octaves,
remap + radial falloff; workload/c2_fastnoise_256.f32 written by tools/make_c2_density.py), sigma_s 1.1, sigma_a 0.01, HG
g=0, density multiplier 100, scale 5 at the origin, point emitter, camera (0,0,-12) -> origin, vfov 45, 1920x1080, 64 spp,
6 bounces. One STEP = one whole frame (132.7 M camera paths) through the wavefront renderer. With N GPUs every rank holds
a scene replica and renders ITS SHARE of the frame's 64 samples per pixel (sample_range(rank, N, 64): STRONG scaling, the
frame is the same at every N); the per-GPU fp32 accumulation buffers are summed onto rank 0 with one NCCL reduce inside
the step. The weak-scaling figure (64 spp per GPU, frame of N*64) is reported beside it under "weak".

Printed JSON (rank 0, one line): see README / the driver contract. `value` = device-timed render (scene resident in
HBM); `e2e` = the same frame through ne_b200_scene_upload + ne_b200_render_frame with HOST buffers (scene H2D, frame
D2H inside the timed region); `roofline` = the volume-tracking kernels (delta tracking `k_wf_track` + ratio
tracking `k_wf_tr`), algorithmic bytes = 40 B per tracking step over their CUDA-event time;
`cpu_baseline` = the reference's own integrator (oracle/_ref, thread-local RNG build) on this box's host cores over a
bounded sample of the same frame.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

W, H, SPP, BOUNCES = 1920, 1080, 64, 6
GRID = 256
METRIC = "Mpaths/s (1080p, 64 spp, heterogeneous 256^3 volume + point light)"
WORKLOAD = "C2: heterogeneous 256^3 density volume, delta tracking + point light, 1920x1080, 64 spp, 6 bounces"


def pinned_like(a):
    """A page-locked host copy of a numpy array (torch's pinned allocator), viewed as numpy: the e2e leg hands the library
    pinned host buffers, as the bench contract asks (host->device copies from pinned host memory)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    v = t.numpy()
    v_keepalive.append(t)
    return v


v_keepalive = []


DENSITY_FILE = os.path.join(ROOT, "workload", f"c2_fastnoise_{GRID}.f32")
DENSITY_SOURCE = "FastNoise SimplexFractal seed 1337, 5 octaves, frequency 4/256 (SURVEY 8d C2; workload/c2_fastnoise_256.f32)"


def load_density():
    """The C2 density: a plain data file (git-ignored, travels with the snapshot; written by tools/make_c2_density.py from
    the reference tree's FastNoise where that tree exists). No oracle code is loaded here."""
    if not (os.path.exists(DENSITY_FILE) and os.path.getsize(DENSITY_FILE) == 4 * GRID ** 3):
        raise SystemExit(f"bench.py: {DENSITY_FILE} is missing: run `python tools/make_c2_density.py` where /root/reference exists "
                         "(__graft_entry__.build() does)")
    return np.fromfile(DENSITY_FILE, np.float32).reshape(GRID, GRID, GRID)


def build_scene(pin=False, transform_fn=None):
    import scenes
    t = time.time()
    grid = load_density()
    if pin:
        grid = pinned_like(grid)
    b = scenes.c2_scene(grid, transform_fn=transform_fn)
    return b, scenes.C2_CAMERA, grid, time.time() - t


def bench_config(world, spp):
    """The `config` object both arms print (identical keys and values for the same N)."""
    return {"workload": WORKLOAD if spp == SPP else WORKLOAD + f" [DEBUG spp={spp}]", "resolution": [W, H], "spp": spp, "bounces": BOUNCES,
            "grid": [GRID] * 3, "density": DENSITY_SOURCE,
            "partition": f"sample-index x{world} (each GPU renders its share of the {spp} spp of every pixel) + NCCL reduce" if world > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write)"}


def csrc_sha16():
    """Hash of the kernel sources: profiles carry it, so a traffic figure measured on other code is refused."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "narvalengine_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h", ".cpp")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every 5 ms from a
    thread (a frame is ~35 ms, too short for `nvidia-smi -lms`); nvidia-smi is the fallback when NVML is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop_flag, self.sm_max = index, [], None, None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self.stop_flag:
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((mhz, [n for n, b in bits.items() if r & b]))
                    except pynvml.NVMLError:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) < 6:
                continue
            try:
                mhz, self.sm_max = float(r[0]), float(r[1])
            except ValueError:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.rows.append((mhz, [n for n, v in zip(names, r[2:6]) if v.lower().startswith("active")]))

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for n in r[1]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def cpu_reference_run(builder, cam, budget_s, threads=None, faithful=False):
    """The reference's own integrator on the host cores over a bounded, frame-covering sample of the workload:
    every `row_step`-th row of the 1080p frame at `spp` samples. Returns (Mpaths/s, description, cores).
    faithful: the build with the reference's RNG exactly as shipped (one racy global mt19937) instead of the
    thread-local patch (SURVEY 8d asks for both; speed-ups are quoted against the faster one)."""
    from refclient import RefOracle
    oracle = RefOracle(faithful=faithful)
    threads = threads or os.cpu_count() or 1
    sc = oracle.scene(builder)
    # pilot to size the sample: every 90th row (12 rows) x 1 spp
    _, secs = sc.render(cam, W, H, 1, BOUNCES, seed=3, threads=threads, row_step=90)
    rate = (len(range(0, H, 90)) * W) / max(secs, 1e-6)
    target_paths = rate * budget_s
    row_step = 18
    rows = len(range(0, H, row_step))
    spp = int(max(1, min(SPP, round(target_paths / (rows * W)))))
    _, secs = sc.render(cam, W, H, spp, BOUNCES, seed=4, threads=threads, row_step=row_step)
    paths = rows * W * spp
    sc.close()
    return paths / secs / 1e6, f"every {row_step}th row of the 1920x1080 frame ({rows} rows) x {spp} spp = {paths} paths in {secs:.1f} s", threads


def run_reference(args, rank):
    if rank != 0:
        return
    if args.cpu_faithful and not args.child:
        # The reference's RNG as shipped is ONE global std::mt19937 shared by all rendering threads (utils/Math.h:59-66): a data
        # race that can corrupt the generator's index and crash. Run that build in a child process; if it dies, fall back to
        # one thread (no race) and say so.
        cmd = [sys.executable, os.path.abspath(__file__)] + [a for a in sys.argv[1:]] + ["--child"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0 and r.stdout.strip():
            print(r.stdout.strip().splitlines()[-1], flush=True)
            return
        r1 = subprocess.run(cmd + ["--ref-threads", "1"], capture_output=True, text=True)
        if r1.returncode == 0 and r1.stdout.strip():
            line = json.loads(r1.stdout.strip().splitlines()[-1])
            line["cpu_baseline"]["note"] = (f"the as-shipped build crashed with all host threads (exit {r.returncode}: racy global mt19937, Math.h:59-66); "
                                            "this is the same build on ONE thread")
            print(json.dumps(line), flush=True)
        else:
            print(json.dumps({"impl": "reference", "unavailable": f"the reference build with its RNG as shipped crashed (exit {r.returncode} / {r1.returncode})"}), flush=True)
        return
    from refclient import RefOracle
    # the scene's transforms come from the oracle itself (getTransform + glm::inverse): the product library is not loaded here
    # (one oracle build per process: the two builds export the same C++ inline variables, which the dynamic linker unifies)
    b, cam, _, _ = build_scene(transform_fn=RefOracle(faithful=args.cpu_faithful).transform_fn())
    vals, ms, sample, cores = [], [], "", 0
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        v, sample, cores = cpu_reference_run(b, cam, budget_s=args.ref_seconds, faithful=args.cpu_faithful, threads=args.ref_threads or None)
        if i >= args.warmup:
            vals.append(v)
            ms.append((time.time() - t0) * 1e3)
    value = float(np.mean(vals))
    kind = "reference"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": bench_config(args.gpus, args.spp),
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample,
                             "rng": "as shipped (one racy global mt19937)" if args.cpu_faithful else "thread_local patch of the global mt19937"},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-seconds", type=float, default=12.0, help="CPU work per reference step / cpu_baseline sample")
    ap.add_argument("--spp", type=int, default=SPP, help="debug only: a value other than 64 is not the headline configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-faithful", action="store_true", help="CPU legs use the reference build with its RNG as shipped (racy global mt19937)")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling measurement at N > 1")
    ap.add_argument("--ref-threads", type=int, default=0, help="host threads of the CPU legs (0 = all)")
    ap.add_argument("--child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from narvalengine_b200.engine import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; narvalengine_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    spp = args.spp
    b, cam_params, grid, t_scene = build_scene(pin=True)  # the host copy of the scene's grid lives in pinned memory
    ctx = Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.upload(b)
    cam = cam_params.make(W / H, ctx.lib)
    ctx.set_camera(cam)
    ctx.render(W, H, 0, 0, BOUNCES)  # allocate the accumulation buffer
    from narvalengine_b200.multigpu import PartitionedFrame, alias_accum
    frame = PartitionedFrame(ctx, alias_accum(ctx, local_rank), rank, world, dist if world > 1 else None)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, n_steps, n_warm):
        """n_warm untimed steps, then exactly n_steps timed ones: L2 flushed and every rank synchronised (barrier +
        cudaDeviceSynchronize) before each, CUDA events on the render stream around each, max over ranks of the sum."""
        for i in range(n_warm):
            step_fn(i)
        sync_all()
        ctx.counters_reset()
        ms = []
        for i in range(n_steps):
            flush.fill_(i & 0xff)  # L2 flush between timed iterations (outside the timed bracket)
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn(n_warm + i)
            e1.record(stream)
            sync_all()
            ms.append(e0.elapsed_time(e1))
        total = float(sum(ms))
        if world > 1:
            t = torch.tensor([total], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, ctx.counters()

    def step(i):
        """One frame: this rank's share of the frame's 64 samples per pixel (strong scaling: the same frame at every N),
        then the one exchange step of the path: the sum of the per-GPU accumulation buffers onto rank 0 (NCCL reduce)."""
        frame.render(W, H, spp, BOUNCES, seed=1 + i)

    def step_weak(i):
        frame.render(W, H, world * spp, BOUNCES, seed=1 + i)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, c = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    paths_per_step = W * H * spp
    value = paths_per_step * args.steps / (total_ms * 1e-3) / 1e6
    gpu_launches = int(c.kernel_launches)
    stage_ms = {"volume": c.ms_volume_kernel, "extend_shadow": c.ms_extend_kernel, "shade": c.ms_shade_kernel, "generate_plan": c.ms_other_kernel,
                "render": c.ms_render}
    counters = {k: int(getattr(c, k)) for k in ("paths", "extend_rays", "shadow_rays", "delta_steps", "ratio_steps", "brick_visits",
                                                  "scatter_events", "wavefront_iterations")}

    weak = None
    if world > 1 and not args.no_weak:
        wk_ms, _ = timed(step_weak, args.steps, 1)
        weak = {"value": W * H * spp * world * args.steps / (wk_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": wk_ms / args.steps,
                "spp_per_gpu": spp, "note": "secondary: 64 spp per GPU, frame of N*64 spp"}

    # ---- roofline of the volume-tracking kernels (this rank). The timed region above runs the production path (one CUDA graph
    # per frame; its per-stage device times are %globaltimer stamps between the stages). The kernel time the roofline uses
    # comes from CUDA EVENTS on the render stream: the same kernels launched by the host-driven loop (NE_B200_HOST_LOOP=1)
    # over the same number of steps of the same workload; the graph's own figure is reported beside it.
    os.environ["NE_B200_HOST_LOOP"] = "1"
    ev_ms, ce = timed(step, args.steps, 1)
    del os.environ["NE_B200_HOST_LOOP"]
    steps_tracked = int(ce.delta_steps + ce.ratio_steps)
    alg_bytes = steps_tracked * int(ce.bytes_per_tracking_step)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ce.ms_volume_kernel * 1e-3) / 1e9 if ce.ms_volume_kernel > 0 else 0.0
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("csrc_sha16") == csrc_sha16():
                traffic = tj.get("dram_bytes_per_launch")
            else:
                traffic_note = f"profiles/r02_traffic.json was measured on other kernel sources ({tj.get('csrc_sha16')} != {csrc_sha16()}): not reported"
        except (OSError, ValueError):
            traffic = None
    launches_volume = max(1, int(ce.wavefront_iterations) * 2)
    frame_ms = ce.ms_volume_kernel + ce.ms_extend_kernel + ce.ms_shade_kernel + ce.ms_other_kernel + 1e-9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_wf_track (delta tracking) + k_wf_tr (ratio tracking)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / launches_volume, "tracking_steps_per_frame": steps_tracked / args.steps,
                "kernel_ms_per_frame": ce.ms_volume_kernel / args.steps, "timer": "CUDA events on the render stream (host-driven loop over the same kernels)",
                "kernel_ms_per_frame_in_graph": c.ms_volume_kernel / args.steps,
                "share_of_step": ce.ms_volume_kernel / frame_ms,
                "note": "these kernels are latency/occupancy-bound, not HBM-bound: see roofline_issue",
                "Mrays_per_s": (c.extend_rays + c.shadow_rays) / (total_ms * 1e-3) / 1e6 * world}
    if traffic_note:
        roofline["traffic_note"] = traffic_note
    # second roofline block for the latency-bound kernels: issue-slot utilisation x lanes per instruction from the committed
    # ncu capture of the same sources (None when the capture belongs to other code)
    roofline_issue = None
    ip = os.path.join(ROOT, "profiles", "r02_issue.json")
    if os.path.exists(ip):
        try:
            ij = json.load(open(ip))
            if ij.get("csrc_sha16") == csrc_sha16():
                roofline_issue = ij.get("roofline_issue")
        except (OSError, ValueError):
            pass

    # ---- e2e: the reference-facing calls with HOST buffers (scene H2D + frame D2H inside the timed region)
    e2e = None
    if (rank == 0 or world > 1) and not args.no_e2e:
        tm = pinned_like(np.empty((H, W, 3), np.float32))  # the frame comes back into pinned host memory
        desc = b.desc()
        h2d = int(grid.nbytes + 4096)
        d2h = int(tm.nbytes)
        ts = []
        for i in range(2 + args.steps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            ctx.upload(desc)
            if world == 1:
                ctx.render_frame(cam, W, H, spp, BOUNCES, 100 + i, 0, tm, None)
            else:
                step(1000 + i)
                if rank == 0:
                    ctx.read_tonemapped(W, H, tm)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ts.append(time.perf_counter() - t0)
        ts = ts[2:]
        e2e_t = float(np.mean(ts))
        if world > 1:
            t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_t = float(t.item())
        e2e = {"value": paths_per_step / e2e_t / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_t * 1e3, "includes": "ne_b200_scene_upload (H2D from pinned host memory + brick build) + ne_b200_render_frame (render, resolve, D2H into pinned host memory)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.cpu_faithful:  # the as-shipped RNG build may crash under threads: measured in a child process (run_reference)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-faithful", "--steps", "1", "--warmup", "0",
                                "--ref-seconds", str(args.ref_seconds)], capture_output=True, text=True)
            try:
                cpu = json.loads(r.stdout.strip().splitlines()[-1]).get("cpu_baseline")
            except (ValueError, IndexError):
                cpu = None
        else:
            v, sample, cores = cpu_reference_run(b, cam_params, budget_s=args.ref_seconds)
            cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "sample": sample,
                   "note": "reference TUs compiled unmodified except a thread_local patch of the global mt19937 (oracle/Makefile)"}

    if rank == 0:
        cfg = bench_config(world, spp)  # the workload, key for key what the reference arm prints
        # what is specific to this arm stays out of `config`: tracking variant, pool size, hash of the kernel sources
        build_info = {"majorant": "per-brick (8^3) DDA", "pool_slots": int(os.environ.get("NE_B200_POOL", 1 << 26)), "csrc_sha16": csrc_sha16()}
        line = {"metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg, "build": build_info,
                "roofline": roofline, "roofline_issue": roofline_issue, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
                "weak": weak, "counters": counters, "kernel_ms": stage_ms}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
