#!/usr/bin/env python
"""Writes workload/c2_fastnoise_<N>.f32: BASELINE configs[1]'s density ("fastnoise-generated 256^3") from
the generator oracle/_ref/libc2noise.so (the reference tree's vendored FastNoise compiled in place + oracle/ref/c2noise.cpp).
Run by __graft_entry__.build() where the generator exists. The data file is git-ignored and travels to the GPU box
with the snapshot (like the built .so files); bench.py and tools/ read it with numpy and never touch oracle/."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def generate(n=256, seed=1337):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libc2noise.so"))
    g = np.zeros((n, n, n), np.float32)
    lib.c2noise_generate(n, n, n, seed, g.ctypes.data_as(C.POINTER(C.c_float)))
    return g


def path(n=256):
    return os.path.join(ROOT, "workload", f"c2_fastnoise_{n}.f32")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    p = path(n)
    if os.path.exists(p) and os.path.getsize(p) == 4 * n ** 3:
        return
    os.makedirs(os.path.dirname(p), exist_ok=True)
    g = generate(n)
    g.tofile(p)
    print(f"{p}: max {g.max():.4f} mean {g.mean():.4f} nonzero {np.count_nonzero(g) / g.size:.3f}")


if __name__ == "__main__":
    main()
