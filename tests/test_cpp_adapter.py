"""The C++ adapter a NarvalEngine maintainer links (include/ne_b200_offline_engine.hpp, OfflineEngine's public surface and
tile protocol over the C ABI): compiled with g++ against the in-tree library; on the CPU it must fail loudly (no
fallback), on a GPU it runs the SceneEditor-style tile loop on a JSON scene and exports PNG/EXR."""
import json
import os
import subprocess

import numpy as np
import pytest

from narvalengine_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = {
    "version": "1",
    "materials": [{"name": "floor", "type": "microfacet", "roughness": 0.95, "metallic": 0.0, "albedo": [.8, .8, .8]},
                  {"name": "light", "type": "emitter", "albedo": [10, 10, 10]}],
    "primitives": [
        {"name": "l", "type": "rectangle", "materialName": "light", "transform": {"position": [0, 2, 0], "scale": [1, 1, 1], "rotation": [-90, 0, 0]}},
        {"name": "f", "type": "rectangle", "materialName": "floor", "transform": {"position": [0, 0, 0], "scale": [6, 6, 1], "rotation": [90, 0, 0]}}],
    "camera": {"position": [0, 1, -3], "lookAt": [0, 0.5, 0], "up": [0, 1, 0], "speed": 5, "vfov": 45, "aperture": 0.0001, "autoFocus": True, "focus": 1},
    "renderer": {"resolution": [120, 60], "spp": 8, "bounces": 6, "mode": "offline", "HDR": False, "toneMapping": False},
}


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "offline_engine_test"
    libdir = os.path.dirname(abi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "offline_engine_test.cpp"),
           "-L", libdir, "-lnarval_b200", f"-Wl,-rpath,{libdir}", "-o", str(out)]
    subprocess.check_call(cmd)
    return str(out)


def _run(binary, tmp_path):
    (tmp_path / "s.json").write_text(json.dumps(SCENE))
    return subprocess.run([binary, str(tmp_path / "s.json"), str(tmp_path), str(tmp_path / "frame")], capture_output=True, text=True, timeout=300)


def test_adapter_compiles_and_fails_loudly_without_a_gpu(binary, tmp_path):
    lib = abi.load_library()
    if lib.ne_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run(binary, tmp_path)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_adapter_tile_protocol_on_the_gpu(binary, tmp_path):
    r = _run(binary, tmp_path)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("OK 120x60 spp 8")
    assert float(r.stdout.split()[-1]) > 0.01
    Image = pytest.importorskip("PIL.Image")
    png = np.asarray(Image.open(tmp_path / "frame.png"))
    assert png.shape == (60, 120, 3) and png.any()
    assert (tmp_path / "frame.exr").stat().st_size > 120 * 60 * 12
