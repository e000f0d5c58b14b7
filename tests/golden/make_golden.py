"""Generates the committed golden fixtures from the ORACLE (the reference's own code, oracle/_ref). Run where
/root/reference exists:  python tests/golden/make_golden.py [case ...]
  image_<case>.npz : two independent converged oracle renders (different mt seeds per thread) of the test_gpu_render
                     cases at equal spp -> parity target + oracle-vs-oracle noise floor.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import scenes  # noqa: E402
from refclient import RefOracle  # noqa: E402

SIZES = {"cornell": (24, 16, 65536), "volume": (24, 16, 65536), "mixed": (16, 12, 524288), "mesh": (24, 16, 131072),
         "directional": (24, 16, 32768), "environment": (24, 16, 32768), "homogeneous": (24, 16, 32768),
         "c2": (48, 27, 16384), "mesh100k": (24, 16, 65536), "pointlit": (24, 16, 32768)}


def main():
    from test_gpu_render import CASES
    o = RefOracle()
    nthreads = os.cpu_count() or 1
    only = [a for a in sys.argv[1:] if a in CASES]
    for name, (mk, cam) in CASES.items():
        if only and name not in only:
            continue
        W, H, spp = SIZES[name]
        sc = o.scene(mk(o.transform_fn()) if name in ("cornell", "mixed") and False else mk())
        t = time.time()
        a, _ = sc.render(cam, W, H, spp, 6, seed=1, threads=nthreads)
        b, _ = sc.render(cam, W, H, spp, 6, seed=2, threads=nthreads)
        np.savez_compressed(os.path.join(HERE, f"image_{name}.npz"), linear=a, linear_b=b, W=W, H=H, spp=spp)
        from imgmetrics import rel_mse
        print(name, W, H, spp, f"{time.time() - t:.1f}s mean {a.mean():.5f} floor {rel_mse(b, a):.3e}")


if __name__ == "__main__":
    main()
