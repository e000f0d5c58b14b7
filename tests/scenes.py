"""Scene fixtures shared by the parity tests, the golden-vector generator and bench.py (synthetic scenes of the
shapes BASELINE.json names)."""
import numpy as np

from narvalengine_b200 import SceneBuilder, CameraParams


def s1_cornell(transform_fn=None, with_sphere=True, li=(400, 400, 400)):
    """S1 of SURVEY Appendix E: the five cornellbox.json walls (resources/scenes/cornellbox.json:107-178), a GGX
    sphere and the rectangle emitter."""
    b = SceneBuilder(transform_fn)
    b.add_microfacet("white", (.8, .8, .8), 0.95, 0.0)
    b.add_microfacet("red", (.8, .02, .05), 0.95, 0.0)
    b.add_microfacet("green", (.11, .8, .01), 0.95, 0.0)
    b.add_microfacet("ball", (.7, .7, .9), 0.5, 0.0)
    b.add_emitter("light", li)
    b.add_rectangle("white", (0, 2, 2), (0, 0, 0), (4, 4, 1))      # back
    b.add_rectangle("white", (0, 0, 0), (90, 0, 0), (4, 4, 1))     # floor
    b.add_rectangle("white", (0, 4, 0), (-90, 0, 0), (4, 4, 1))    # ceiling
    b.add_rectangle("red", (-2, 2, 0), (0, -90, 0), (4, 4, 1))     # left
    b.add_rectangle("green", (2, 2, 0), (0, 90, 0), (4, 4, 1))     # right
    if with_sphere:
        b.add_sphere("ball", (.6, .6, .5), 0.6)
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    return b


CORNELL_CAMERA = CameraParams((0, 2, -5), (0, 2, 0), 45.0)


def s2_grid():
    g = np.zeros((4, 4, 4), np.float32)
    for z in range(4):
        for y in range(4):
            for x in range(4):
                g[z, y, x] = (x + 2 * y + 3 * z) / 18.0
    return g


def s2_volume(transform_fn=None, leaves=False):
    """S2 of SURVEY Appendix E: 4^3 grid volume at (0,1,0) scale 2 + the S1 light."""
    b = SceneBuilder(transform_fn)
    if leaves:
        v = np.zeros((1, 8, 8, 8), np.float32)
        v[0, :4, :4, :4] = s2_grid()
        vol = b.add_volume_leaves((4, 4, 4), [[0, 0, 0]], v)
    else:
        vol = b.add_volume_dense(s2_grid())
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 3.0, vol, "hg", 0.0)
    b.add_emitter("light", (400, 400, 400))
    b.add_volume("cloud", (0, 1, 0), (0, 0, 0), (2, 2, 2))
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    return b
