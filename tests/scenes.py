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


# ---------------------------------------------------------------------------------------------------------------
# Synthetic data of the shapes BASELINE.json names
# ---------------------------------------------------------------------------------------------------------------
def fbm_noise(res, seed=1337, base=4, octaves=5):
    """Deterministic value-noise fBm in [0,1], array indexed [z,y,x]. (BASELINE config 2 says "fastnoise-generated";
    FastNoise lives in the reference's includes/ and is unused by its hot path, so the density field is generated
    here with the same character: 5 octaves, base frequency 4 cells across the volume.)"""
    from scipy.ndimage import zoom
    W, H, D = res
    rng = np.random.default_rng(seed)
    out = np.zeros((D, H, W), np.float32)
    amp, tot, f = 1.0, 0.0, base
    for _ in range(octaves):
        lat = rng.uniform(0, 1, (f + 1, f + 1, f + 1)).astype(np.float32)
        up = zoom(lat, (D / (f + 1), H / (f + 1), W / (f + 1)), order=1, mode="nearest", grid_mode=True)
        out += amp * up[:D, :H, :W]
        tot += amp
        amp *= 0.5
        f *= 2
    return out / tot


def cloud_density(res, seed=1337, threshold=0.45):
    """fBm remapped to a sparse density in [0,1] with a radial falloff (config 2 shape: dense core, empty rim)."""
    W, H, D = res
    n = fbm_noise(res, seed)
    z, y, x = np.meshgrid(np.linspace(-.5, .5, D), np.linspace(-.5, .5, H), np.linspace(-.5, .5, W), indexing="ij")
    r = np.sqrt(x * x + y * y + z * z)
    t = np.clip((0.5 - r) / 0.15, 0, 1)
    fall = t * t * (3 - 2 * t)
    g = np.maximum(0, n - threshold) / (1 - threshold) * fall
    g = g / max(g.max(), 1e-9)
    return g.astype(np.float32)


def dense_to_leaves(grid):
    """Split a [z,y,x] grid into the 8^3 leaves an OpenVDB leaf iterator would yield (inactive = all-zero omitted)."""
    D, H, W = grid.shape
    pd, ph, pw = (-D) % 8, (-H) % 8, (-W) % 8
    g = np.pad(grid, ((0, pd), (0, ph), (0, pw)))
    bz, by, bx = g.shape[0] // 8, g.shape[1] // 8, g.shape[2] // 8
    blocks = g.reshape(bz, 8, by, 8, bx, 8).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 8, 8, 8)
    iz, iy, ix = np.meshgrid(np.arange(bz), np.arange(by), np.arange(bx), indexing="ij")
    origins = np.stack([ix.ravel() * 8, iy.ravel() * 8, iz.ravel() * 8], 1).astype(np.int32)
    act = blocks.reshape(len(blocks), -1).any(axis=1)
    return origins[act], np.ascontiguousarray(blocks[act])


def s2_volume(transform_fn=None, leaves=False, phase="hg", g=0.0):  # noqa: F811  (extends the definition above)
    b = SceneBuilder(transform_fn)
    if leaves:
        v = np.zeros((1, 8, 8, 8), np.float32)
        v[0, :4, :4, :4] = s2_grid()
        vol = b.add_volume_leaves((4, 4, 4), [[0, 0, 0]], v)
    else:
        vol = b.add_volume_dense(s2_grid())
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 3.0, vol, phase, g)
    b.add_emitter("light", (400, 400, 400))
    b.add_volume("cloud", (0, 1, 0), (0, 0, 0), (2, 2, 2))
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    return b


def noise_volume_scene(res=(64, 64, 64), leaves=False, density=100.0, light="point", scale=(2.4, 2.4, 2.4), pos=(0, 1, 0),
                       li=(100, 100, 70), transform_fn=None, grid=None, g=0.0):
    """Config-2 shape (resources/scenes/volumes.json): one heterogeneous grid volume + one emitter."""
    b = SceneBuilder(transform_fn)
    grid = cloud_density(res) if grid is None else grid
    if leaves:
        o, v = dense_to_leaves(grid)
        vol = b.add_volume_leaves(res, o, v)
    else:
        vol = b.add_volume_dense(grid)
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), density, vol, "hg", g)
    b.add_emitter("light", li)
    b.add_volume("cloud", pos, (0, 0, 0), scale)
    if light == "point":
        b.add_point("light", (0, 6, 0))
    else:
        b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1.5, 1.5, 1))
    return b


def c2_density(n=256, seed=1337):
    """BASELINE configs[1] density as SURVEY 8d C2 specifies it: FastNoise SimplexFractal (seed 1337, frequency 4/N, 5
    octaves) remapped max(0, v*0.5+0.5-0.35)/0.65, times smoothstep(0.5, 0.35, |p|). Read from the data file
    tools/make_c2_density.py wrote (workload/, travels to the GPU box), else generated by oracle/_ref/libc2noise.so
    (the reference tree's vendored FastNoise compiled in place, oracle/ref/c2noise.cpp)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(root, "workload", f"c2_fastnoise_{n}.f32")
    if seed == 1337 and os.path.exists(p) and os.path.getsize(p) == 4 * n ** 3:
        return np.fromfile(p, np.float32).reshape(n, n, n)
    import ctypes as C
    lib = C.CDLL(os.path.join(root, "oracle", "_ref", "libc2noise.so"))
    g = np.zeros((n, n, n), np.float32)
    lib.c2noise_generate(n, n, n, seed, g.ctypes.data_as(C.POINTER(C.c_float)))
    return g


def c2_scene(grid=None, n=256, transform_fn=None):
    """BASELINE configs[1] / SURVEY 8d C2, exactly what bench.py renders: FastNoise density, sigma_s 1.1, sigma_a 0.01,
    HG g=0, density multiplier 100, volume scale 5 at the origin, POINT emitter Li=(100,100,70) at (0,6,0)."""
    grid = c2_density(n) if grid is None else grid
    return noise_volume_scene(res=grid.shape[::-1], density=100.0, light="point", scale=(5, 5, 5), pos=(0, 0, 0), li=(100, 100, 70),
                              grid=grid, transform_fn=transform_fn)


C2_CAMERA = CameraParams((0, 0, -12), (0, 0, 0), 45.0)


def point_lit_surface_scene(transform_fn=None):
    """A GGX floor and back wall lit by a POINT emitter only (primitives/Point.cpp:10-26; position = M * p with M already
    translated by p, Q25): the light of BASELINE configs[1] on surfaces."""
    b = SceneBuilder(transform_fn)
    b.add_microfacet("floor", (.8, .7, .6), 0.6, 0.1)
    b.add_microfacet("wall", (.3, .5, .8), 0.9, 0.0)
    b.add_emitter("light", (100, 100, 70))
    b.add_rectangle("floor", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    b.add_rectangle("wall", (0, 2, 3), (0, 0, 0), (8, 6, 1))
    b.add_point("light", (0.3, 1.5, 0.2))
    return b


def displaced_grid(n, size=4.0, amp=0.35, seed=11):
    """n x n quads (2 n^2 triangles) displaced by fBm: the config-4 mesh shape at any size."""
    from scipy.ndimage import zoom
    rng = np.random.default_rng(seed)
    h = np.zeros((n + 1, n + 1), np.float32)
    a, f = 1.0, 4
    while f <= max(4, n):
        lat = rng.uniform(-1, 1, (f + 1, f + 1)).astype(np.float32)
        h += a * zoom(lat, ((n + 1) / (f + 1),) * 2, order=1, mode="nearest", grid_mode=True)[:n + 1, :n + 1]
        a *= 0.5
        f *= 2
        if a < 0.02:
            break
    xs = np.linspace(-size / 2, size / 2, n + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    pos = np.stack([X, amp * h, Z], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([(X / size + .5), (Z / size + .5)], -1).reshape(-1, 2).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + n + 1, v00 + n + 2
    # wound so that the geometric normal cross(v1-v0, v2-v0) points up (+y)
    idx = np.stack([np.stack([v00, v01, v10], 1), np.stack([v10, v01, v11], 1)], 1).reshape(-1, 3).astype(np.uint32)
    return pos, idx, uv


def mesh_scene(n=64, transform_fn=None, roughness=0.6, li=(60, 60, 60), normal_map=None, normal_image=None):
    """Config-4 shape: a displaced-grid OBJ-style mesh with one GGX material and two rectangle area lights. normal_map /
    normal_image: the material carries TextureName::NORMAL, so Triangle::intersect bends the geometric normal
    (primitives/Triangle.cpp:74-77, utils/Math.h:1209-1215)."""
    b = SceneBuilder(transform_fn)
    b.add_microfacet("mesh", (.7, .6, .5), roughness, 0.0, normal_map=normal_map, normal_image=normal_image)
    b.add_emitter("light", li)
    pos, idx, uv = displaced_grid(n)
    b.add_mesh("mesh", pos, idx, uv, pos=(0, 0, 0))
    b.add_rectangle("light", (-1, 3.5, 0), (-90, 0, 0), (1.5, 1.5, 1))
    b.add_rectangle("light", (1.5, 3.0, 1), (-70, 0, 20), (1, 1, 1))
    return b


MESH_CAMERA = CameraParams((0, 3.2, -4.2), (0, 0, 0), 45.0)


def mixed_scene(sort_and_group=True, transform_fn=None, res=(32, 24, 40), mesh_n=24):
    """Config-5 shape: mesh + rectangles + sphere light + sparse cloud volume (the volume deliberately FIRST in JSON
    order so sort_and_group changes the fold)."""
    b = SceneBuilder(transform_fn, sort_and_group=sort_and_group)
    o, v = dense_to_leaves(cloud_density(res, seed=7))
    vol = b.add_volume_leaves(res, o, v)
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 25.0, vol, "hg", 0.0)
    b.add_microfacet("floor", (.8, .8, .8), 0.95, 0.0)
    b.add_microfacet("mesh", (.7, .5, .3), 0.8, 0.1)
    b.add_emitter("light", (300, 300, 260))
    b.add_emitter("bulb", (800, 700, 500))
    b.add_volume("cloud", (0, 1.6, 0.5), (0, 20, 0), (2.2, 1.4, 1.8))
    b.add_rectangle("floor", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    b.add_rectangle("floor", (0, 2, 3), (0, 0, 0), (8, 6, 1))
    pos, idx, uv = displaced_grid(mesh_n, size=3.0, amp=0.25)
    b.add_mesh("mesh", pos, idx, uv, pos=(0, 0.3, 0))
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    b.add_sphere("bulb", (2.3, 2.5, 0), 0.05)
    return b


MIXED_CAMERA = CameraParams((0, 2, -5), (0, 1.2, 0), 45.0)


def textured_scene(transform_fn=None):
    """A floor with an RGBA8 image albedo (mirror wrap, nearest texel) like cornellbox.json's checkboard.png."""
    b = SceneBuilder(transform_fn)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (16, 24, 4), dtype=np.uint8)
    b.add_microfacet("tex", None, 0.7, 0.1, albedo_image=img)
    b.add_emitter("light", (50, 50, 50))
    b.add_rectangle("tex", (0, 0, 0), (90, 0, 0), (4, 4, 1))
    b.add_rectangle("light", (0, 3, 0), (-90, 0, 0), (1, 1, 1))
    return b


def cornell_c1(transform_fn=None):
    """BASELINE config 1: Cornell-box-style scene (cornellbox.json walls, GGX) + area emitter + a sphere. The sphere is
    an EMITTER (as in resources/scenes/playground.json): the reference never terminates on a GGX sphere
    (DESIGN.md "Deviations": intersectTr loops forever on the 'inside' self-hit, Model.cpp:418-424)."""
    b = s1_cornell(transform_fn, with_sphere=False)
    b.add_emitter("bulb", (200, 160, 120))
    b.add_sphere("bulb", (.9, .8, .3), 0.15)
    return b


def directional_scene(transform_fn=None, res=(20, 16, 24), g=0.2):
    """A `directionalLight` sun (carried by a point primitive) over a cloud and a floor, plus a small area light: the sun
    reaches the image only through Light::Le on camera rays and through the BSDF half of estimateDirect (Q23, Q12)."""
    b = SceneBuilder(transform_fn)
    vol = b.add_volume_dense(cloud_density(res, seed=5))
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 20.0, vol, "hg", g)
    b.add_microfacet("floor", (.8, .8, .8), 0.95, 0.0)
    b.add_directional("sun", (0.6, 0.5, 0.4), (3, 5, -2))
    b.add_emitter("light", (200, 200, 180))
    b.add_volume("cloud", (0, 1.2, 0), (0, 15, 0), (2.2, 1.6, 2.0))
    b.add_rectangle("floor", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    b.add_point("sun", (3, 5, -2))
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    return b


DIRECTIONAL_CAMERA = CameraParams((0, 2, -5), (0, 1, 0), 45.0)


def environment_scene(transform_fn=None, res=(16, 12, 20)):
    """An `infiniteAreaLight` lat-long map carried by a big sphere (the reference casts its primitive to Sphere,
    InfiniteAreaLight.h:95) over a floor, a cloud and a small area light."""
    b = SceneBuilder(transform_fn)
    rng = np.random.default_rng(17)
    env = rng.integers(40, 256, (8, 16, 4), dtype=np.uint8)  # no black texel: Le never is, so the sphere is never hit
    vol = b.add_volume_dense(cloud_density(res, seed=9))
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 15.0, vol, "hg", 0.0)
    b.add_microfacet("floor", (.8, .8, .8), 0.95, 0.0)
    b.add_infinite_area_light("sky", env)
    b.add_emitter("light", (150, 150, 140))
    b.add_volume("cloud", (0.5, 1.2, 0), (0, 10, 0), (1.8, 1.4, 1.8))
    b.add_rectangle("floor", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    b.add_sphere("sky", (0, 0, 0), 40.0)
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    return b


ENVIRONMENT_CAMERA = CameraParams((0, 2, -5), (0, 1, 0), 45.0)


def homogeneous_scene(transform_fn=None, grey=True, g=0.3):
    """A HomogeneousMedia box (volume material without a grid, materials/HomogeneousMedia.cpp) over a floor with an area
    light. (The reference's own SceneReader cannot build this - DESIGN.md §8 - but its integrator handles it.)"""
    b = SceneBuilder(transform_fn)
    sc = (1.1, 1.1, 1.1) if grey else (1.2, 0.9, 0.6)
    b.add_volume_material("fog", sc, (.01, .01, .01), 1.5, -1, "hg", g)
    b.add_microfacet("floor", (.8, .8, .8), 0.95, 0.0)
    b.add_emitter("light", (200, 200, 180))
    b.add_volume("fog", (0.2, 1.1, 0.3), (0, 25, 0), (1.6, 1.2, 1.4))
    b.add_rectangle("floor", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    return b


HOMOGENEOUS_CAMERA = CameraParams((0, 2, -5), (0, 1, 0), 45.0)
