"""Golden fixture for a scene the reference SHIPS: resources/scenes/testing/surfaceAndLight.json, loaded by the library's own
front end (ne_b200_scene_file_load), rendered by the ORACLE (the reference's integrator) twice with different seeds at
the file's own resolution scaled down (150x75, aspect kept) and a converged sample count. The JSON text travels inside
the .npz so that the GPU box (where /root/reference does not exist) loads exactly the shipped file.
Run where /root/reference exists:  python tests/golden/make_scene_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from imgmetrics import rel_mse  # noqa: E402
from narvalengine_b200.scene import SceneFile, CameraParams  # noqa: E402
from refclient import RefOracle  # noqa: E402

PATH = "/root/reference/resources/scenes/testing/surfaceAndLight.json"
W, H, SPP = 48, 24, 16384


def main():
    text = open(PATH).read()
    sf = SceneFile(text=text, resources_dir="/root/reference/resources/")
    o = RefOracle()
    sc = o.scene(sf.desc())
    # camera of the file (SceneReader.cpp:650-668: aperture forced to 1e-4, autoFocus -> focus 3)
    cam = CameraParams((0, 1, -3), (0, 2, 0), 45.0)
    st = sf.settings()
    a, _ = sc.render(cam, W, H, SPP, st.bounces, seed=1, threads=os.cpu_count() or 1)
    b, _ = sc.render(cam, W, H, SPP, st.bounces, seed=2, threads=os.cpu_count() or 1)
    np.savez_compressed(os.path.join(HERE, "scene_surfaceAndLight.npz"), scene_json=np.frombuffer(text.encode(), np.uint8), linear=a, linear_b=b,
                        W=W, H=H, spp=SPP, bounces=st.bounces)
    print(f"surfaceAndLight {W}x{H}x{SPP}: mean {a.mean():.5f} floor {rel_mse(b, a):.3e}")


if __name__ == "__main__":
    main()
