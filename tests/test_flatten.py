"""The compiled reference-side adapter (oracle/ref/flatten.cpp: narvalengine::Scene* -> ne_b200_scene_desc, Camera ->
ne_b200_camera), round-trip tested: descriptor A -> the reference's object graph (the harness follows
SceneReader.cpp:67-675) -> flatten -> descriptor B. B must describe the same scene as A: same primitives in fold order
(models, then lights), same transforms bit for bit, same material parameters and texture / grid / mesh contents. The GPU
half renders A and B and compares the images."""
import ctypes as C

import numpy as np
import pytest

import scenes
from narvalengine_b200 import abi
from refclient import RefOracle


@pytest.fixture(scope="module")
def oracle():
    return RefOracle()


SCENES = {
    "cornell": scenes.s1_cornell,
    "mixed_sorted": lambda tf=None: scenes.mixed_scene(sort_and_group=True, transform_fn=tf),
    "mixed_json_order": lambda tf=None: scenes.mixed_scene(sort_and_group=False, transform_fn=tf),
    "point_lit_volume": lambda tf=None: scenes.noise_volume_scene(res=(24, 20, 16), density=12.0, light="point", transform_fn=tf),
    "textured": scenes.textured_scene,
    "normal_mapped_mesh": lambda tf=None: scenes.mesh_scene(n=12, normal_map=(0.35, 0.6, 0.95), transform_fn=tf),
    "directional": scenes.directional_scene,
    "environment": scenes.environment_scene,
    "homogeneous": scenes.homogeneous_scene,
}


def fold_order(d):
    """Primitive indices of a descriptor in the reference's fold order: non-emitters (media last when sort_and_group), emitters."""
    prims = [d.primitives[i] for i in range(d.n_primitives)]
    mats = [d.materials[i] for i in range(d.n_materials)]

    def is_light(p):
        return p.material >= 0 and mats[p.material].type in (abi.MAT_EMITTER, abi.MAT_DIRECTIONAL, abi.MAT_INFINITE) and p.type != abi.PRIM_MESH \
            and p.type != abi.PRIM_VOLUME

    models = [i for i, p in enumerate(prims) if not is_light(p)]
    lights = [i for i, p in enumerate(prims) if is_light(p)]
    if d.sort_and_group:
        med = [i for i in models if prims[i].material >= 0 and mats[prims[i].material].type == abi.MAT_VOLUME]
        models = [i for i in models if i not in med] + med
    return models + lights


def tex_bytes(t):
    n = {abi.TEX_R32F: 4, abi.TEX_RG32F: 8, abi.TEX_RGB32F: 12, abi.TEX_RGBA32F: 16, abi.TEX_RGBA8: 4}[t.format] * t.width * t.height
    return (t.width, t.height, t.format, t.wrap_u, t.wrap_v, C.string_at(t.texels, n))


def mat_key(d, mi):
    if mi < 0:
        return None
    m = d.materials[mi]
    tex = lambda i: tex_bytes(d.textures[i]) if i >= 0 else None  # noqa: E731
    k = [m.type]
    if m.type == abi.MAT_MICROFACET:
        k += [tex(m.albedo_tex), tex(m.metallic_tex), tex(m.roughness_tex), tex(m.normal_tex), m.has_normal_flag]
    elif m.type == abi.MAT_EMITTER:
        k += [tuple(m.li)]
    elif m.type == abi.MAT_DIRECTIONAL:
        k += [tuple(m.li), tuple(m.direction)]
    elif m.type == abi.MAT_INFINITE:
        k += [tex(m.env_tex)]
    else:
        k += [tuple(m.scattering), tuple(m.absorption), m.density_multiplier, m.phase, m.g if m.phase == abi.PHASE_HG else 0.0]
        if m.volume >= 0:
            v = d.volumes[m.volume]
            if v.dense:
                grid = np.ctypeslib.as_array(v.dense, shape=(v.depth, v.height, v.width)).copy()
            else:  # leaves -> the dense grid tools::copyToDense would leave
                grid = np.zeros((v.depth, v.height, v.width), np.float32)
                o = np.ctypeslib.as_array(v.leaf_origin, shape=(v.n_leaves, 3))
                val = np.ctypeslib.as_array(v.leaf_values, shape=(v.n_leaves, 8, 8, 8))
                for (x, y, z), b in zip(o, val):
                    sub = grid[z:z + 8, y:y + 8, x:x + 8]
                    sub[...] = b[:sub.shape[0], :sub.shape[1], :sub.shape[2]]
            k += [(v.width, v.height, v.depth), grid.tobytes()]
        else:
            k += [None]
    return k


def prim_key(d, i):
    p = d.primitives[i]
    k = [p.type, tuple(p.to_world), tuple(p.to_object), p.collision, mat_key(d, p.material)]
    if p.type == abi.PRIM_SPHERE:
        k.append(p.radius)
    if p.type == abi.PRIM_POINT:
        k.append(tuple(p.point))
    if p.type == abi.PRIM_MESH:
        # the reference's BVH build reorders Model::primitives, so the triangles come back in another ORDER (each with its own
        # three indices in their original order, which is what Q30's uv lookup depends on): compare them as a sorted list
        tris = np.ctypeslib.as_array(p.indices, shape=(p.n_triangles, 3))
        tris = tris[np.lexsort((tris[:, 2], tris[:, 1], tris[:, 0]))]
        k += [p.n_vertices, p.n_triangles, C.string_at(p.positions, 12 * p.n_vertices), tris.tobytes(),
              C.string_at(p.uvs, 8 * p.n_vertices) if p.uvs else None]
    return k


@pytest.mark.parametrize("name", list(SCENES))
def test_flatten_round_trip(oracle, name):
    b = SCENES[name](oracle.transform_fn())
    A = b.desc()
    rs = oracle.scene(A)
    flat = rs.flatten()
    B = flat.desc()
    assert B.n_primitives == A.n_primitives and B.sort_and_group == 0
    order = fold_order(A)
    assert rs.counts() == (sum(1 for _ in order) - sum(1 for i in order[len(order) - rs.counts()[1]:]), rs.counts()[1])
    for j, i in enumerate(order):
        ka, kb = prim_key(A, i), prim_key(B, j)
        if A.primitives[i].type == abi.PRIM_MESH and not A.primitives[i].uvs:
            ka[-1] = kb[-1]  # a mesh without uvs comes back with the zeros assimp-style import stores
        assert ka == kb, f"{name}: primitive {i} (fold position {j}) differs after the round trip"
    # and the flattened descriptor builds the same reference scene again: fixed rays, identical hits
    rs2 = oracle.scene(B)
    rng = np.random.default_rng(3)
    o = (np.array([0, 2, -3]) + rng.uniform(-1, 1, (2000, 3))).astype(np.float32)
    d = rng.normal(size=(2000, 3)).astype(np.float32)
    h1, h2 = rs.intersect(o, d), rs2.intersect(o, d)
    for k in range(2000):
        assert (h1[k].hit, h1[k].instance, h1[k].t_near, tuple(h1[k].normal)) == (h2[k].hit, h2[k].instance, h2[k].t_near, tuple(h2[k].normal))
    flat.close()


def test_camera_to_pod_is_what_the_library_builds(oracle):
    """toPod(Camera) of the adapter (through neref_camera_make) against ne_b200_camera_make, bit for bit."""
    lib = abi.load_library()
    for cp, aspect in ((scenes.CORNELL_CAMERA, 1.0), (scenes.C2_CAMERA, 1920 / 1080), (scenes.MESH_CAMERA, 4 / 3)):
        a, r = cp.make(aspect, lib), oracle.camera_make(cp, aspect)
        for f in ("position", "lower_left", "horizontal", "vertical", "side", "up"):
            assert tuple(getattr(a, f)) == tuple(getattr(r, f)), f
        assert a.lens_radius == r.lens_radius


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mixed_json_order", "point_lit_volume", "normal_mapped_mesh", "environment"])
def test_flattened_scene_renders_like_the_original(oracle, name, monkeypatch):
    from narvalengine_b200.engine import Context
    monkeypatch.setenv("NE_B200_TRACK_BUDGET", "100000000")
    b = SCENES[name]()
    rs = oracle.scene(b)
    flat = rs.flatten()
    ctx = Context(0)
    W, H = 64, 48
    cam = scenes.CameraParams((0, 2, -5), (0, 1, 0), 45.0).make(W / H, ctx.lib)
    imgs = []
    for scene in (b, flat.desc()):
        ctx.upload(scene)
        lin = np.zeros((H, W, 3), np.float32)
        ctx.render_frame(cam, W, H, 16, 6, 5, 0, None, lin)
        imgs.append(lin)
    assert imgs[0].mean() > 0
    np.testing.assert_allclose(imgs[1], imgs[0], rtol=2e-4, atol=1e-5 * float(imgs[0].mean()))
    ctx.close()
    flat.close()
