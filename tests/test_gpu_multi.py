"""The single-process multi-GPU path behind the C ABI (ne_b200_create_multi, csrc/ne_multi.cu): scene
replicas, sample-index partition, worker thread per device, fused peer-memory reduce + resolve. On a one-GPU box the
device list names GPU 0 twice (two contexts, same code path); with two or more GPUs the real peers are used as well."""
import json
import os
import subprocess

import numpy as np
import pytest

import scenes
from narvalengine_b200 import abi
from narvalengine_b200.engine import Context, MultiContext
from narvalengine_b200.multigpu import sample_range

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_lists():
    n = abi.load_library().ne_b200_device_count()
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    return out


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "gpus" + "".join(map(str, d)))
def test_multi_gpu_frame_equals_single_gpu_frame(devices, monkeypatch):
    monkeypatch.setenv("NE_B200_TRACK_BUDGET", "100000000")  # uncut walks: identical paths whatever the partition
    b = scenes.mixed_scene()
    W, H, spp = 96, 64, 10  # 10 samples over 3 ranks: uneven shares
    one = Context(0)
    one.upload(b)
    cam = scenes.MIXED_CAMERA.make(W / H, one.lib)
    ref_lin, ref_tm = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
    one.render_frame(cam, W, H, spp, 6, 7, 0, ref_tm, ref_lin)
    one.counters_reset()
    one.render_frame(cam, W, H, spp, 6, 7, 0, None, None)
    c1 = one.counters()
    m = MultiContext(devices)
    assert len(m) == len(devices)
    m.upload(b)
    for r in range(len(devices)):
        m.rank(r).counters_reset()
    lin, tm = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
    m.render_frame(cam, W, H, spp, 6, 7, 0, tm, lin)
    np.testing.assert_allclose(lin, ref_lin, rtol=2e-4, atol=1e-5 * float(ref_lin.mean()))
    np.testing.assert_allclose(tm, ref_tm, rtol=2e-4, atol=1e-5)
    # every rank traced exactly its share of the paths, and together they traced what one GPU traces
    per_rank = [m.rank(r).counters() for r in range(len(devices))]
    for r, c in enumerate(per_rank):
        b0, b1 = sample_range(r, len(devices), spp)
        assert c.paths == W * H * (b1 - b0)
    for k in ("paths", "extend_rays", "shadow_rays", "scatter_events", "surface_events", "tri_tests"):
        assert sum(int(getattr(c, k)) for c in per_rank) == int(getattr(c1, k)), k
    # rank 0's accumulation buffer holds the whole frame afterwards (checkpointing)
    sums, n = m.rank(0).accum_download(W, H)
    assert n == spp
    np.testing.assert_allclose(sums / spp, ref_lin, rtol=2e-4, atol=1e-5 * float(ref_lin.mean()))
    m.close()
    one.close()


def test_multi_gpu_render_is_asynchronous_and_reusable():
    """ne_b200_multi_render returns while the GPUs work; ne_b200_multi_resolve joins. A second frame reuses everything."""
    b = scenes.noise_volume_scene(res=(32, 32, 32), density=30.0, light="point")
    m = MultiContext([0, 0])
    m.upload(b)
    W, H = 64, 48
    cam = scenes.CameraParams((0, 1, -6), (0, 1, 0), 45.0).make(W / H, m.lib)
    a, a2 = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
    m.render(cam, W, H, 8, 6, seed=3)
    m.resolve(None, a)
    m.render(cam, W, H, 8, 6, seed=3)
    m.resolve(None, a2)
    assert a.mean() > 0 and np.isfinite(a).all()
    np.testing.assert_allclose(a, a2, rtol=2e-4, atol=1e-6)
    m.close()


def test_multi_gpu_argument_errors():
    lib = abi.load_library()
    with pytest.raises(abi.NarvalB200Error):
        MultiContext([])
    with pytest.raises(abi.NarvalB200Error):
        MultiContext([lib.ne_b200_device_count() + 3])


def test_cpp_multi_gpu_program(tmp_path):
    """tests/cpp/multi_gpu_test.cpp: B200OfflineEngine with a device list, compiled C++ through the C ABI."""
    from test_cpp_adapter import SCENE
    out = tmp_path / "multi_gpu_test"
    libdir = os.path.dirname(abi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "multi_gpu_test.cpp"),
                           "-L", libdir, "-lnarval_b200", f"-Wl,-rpath,{libdir}", "-o", str(out)])
    (tmp_path / "s.json").write_text(json.dumps(SCENE))
    n = abi.load_library().ne_b200_device_count()
    devs = "0,1" if n >= 2 else "0,0"
    r = subprocess.run([str(out), str(tmp_path / "s.json"), str(tmp_path), devs], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert r.stdout.startswith("OK 2 devices 120x60 spp 8")
