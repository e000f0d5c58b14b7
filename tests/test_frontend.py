"""Scene front end and framebuffer consumers (csrc/ne_frontend.cpp) on the CPU: the JSON loader must
produce the descriptor that SceneReader::processMaterial / processPrimitives / processCameraAndRenderer imply (checked
against the Python SceneBuilder, itself checked against the reference-built oracle in test_host.py / test_gpu_*), the
.vol reader must reproduce ResourceManager::loadVolasTexture's token rules, and the image writers saveImage's pixels."""
import ctypes as C
import json
import os
import struct
import zlib

import numpy as np
import pytest

pytest.importorskip("PIL")

import scenes
from narvalengine_b200 import abi
from narvalengine_b200.scene import SceneBuilder, SceneFile, read_vol, save_image, write_vol


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()


SCENE = {
    "version": "0.0.1a",
    "comment": "unknown keys are ignored",
    "materials": [
        {"name": "white", "type": "microfacet", "roughness": 0.95, "metallic": 0.0, "albedo": [0.8, 0.8, 0.8]},
        {"name": "tex", "type": "microfacet", "roughness": 0.5, "metallic": 0.25, "albedo": "imgs/check.png", "normalMap": [0.5, 0.5, 1.0]},
        {"name": "emitter", "type": "emitter", "albedo": [10, 9, 8]},
        {"name": "cloud", "type": "volume", "scattering": [1.1, 1.1, 1.1], "absorption": [0.01, 0.01, 0.01], "phaseFunction": "hg", "g": 0.3,
         "density": 25, "path": "vol/small.vol"},
        {"name": "white", "type": "microfacet", "roughness": 0.3, "metallic": 1.0, "albedo": [0.1, 0.2, 0.3]},
    ],
    "primitives": [
        {"name": "vol", "type": "volume", "transform": {"position": [0, 1.5, 0.5], "scale": [2, 1.5, 2], "rotation": [0, 20, 0]}, "materialName": "cloud"},
        {"name": "floor", "type": "rectangle", "transform": {"position": [0, 0, 0], "scale": [8, 8, 1], "rotation": [90, 0, 0]}, "materialName": "white"},
        {"name": "wall", "type": "rectangle", "transform": {"position": [0, 2, 3], "scale": [8, 6, 1], "rotation": [0, 0, 0]}, "materialName": "tex"},
        {"name": "mesh", "type": "obj", "path": "models/quad.obj", "materialName": "tex",
         "transform": {"position": [0, 0.3, 0], "scale": [1, 2, 1], "rotation": [10, 20, 30]}},
        {"name": "light", "type": "rectangle", "transform": {"position": [0, 3.9, 0], "scale": [1, 1, 1], "rotation": [-89, 0, 0]}, "materialName": "emitter"},
        {"name": "bulb", "type": "sphere", "radius": 0.05, "collision": False,
         "transform": {"position": [2.3, 2.5, 0], "scale": [5, 5, 5], "rotation": [45, 0, 0]}, "materialName": "emitter"},
        {"name": "pt", "type": "point", "transform": {"position": [0, 6, 0], "scale": [1, 1, 1], "rotation": [0, 0, 0]}, "materialName": "emitter"},
    ],
    "camera": {"position": [0, 2, -5], "lookAt": [0, 1.2, 0], "up": [0, 0, 1], "speed": 5, "vfov": 40, "aperture": 0.5, "autoFocus": True, "focus": 1},
    "renderer": {"resolution": [96, 64], "spp": 8, "bounces": 6, "mode": "offline", "HDR": False, "toneMapping": False},
}
OBJ = """# a quad and a triangle, negative indices, v/vt/vn forms
v -1 0 -1
v 1 0 -1
v 1 0.5 1
v -1 0 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 1 0
f 1/1/1 2/2/1 3/3/1 4/4/1
f -4//1 -3//1 -1//1
"""


@pytest.fixture(scope="module")
def resources(tmp_path_factory, lib):
    from PIL import Image
    root = tmp_path_factory.mktemp("resources")
    for d in ("imgs", "vol", "models", "scenes"):
        (root / d).mkdir()
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (6, 10, 4), dtype=np.uint8)
    Image.fromarray(img, "RGBA").save(root / "imgs" / "check.png")
    grid = rng.random((5, 4, 6), dtype=np.float32)  # [z, y, x]
    grid[grid < 0.3] = 0
    write_vol(root / "vol" / "small.vol", grid, lib)
    (root / "models" / "quad.obj").write_text(OBJ)
    (root / "scenes" / "scene.json").write_text(json.dumps(SCENE, indent=1))
    return root, img, grid


def tex_value(desc, idx):
    t = desc.textures[idx]
    if t.format == abi.TEX_RGBA8:
        return np.ctypeslib.as_array(C.cast(t.texels, C.POINTER(C.c_uint8)), (t.height, t.width, 4)).copy(), (t.wrap_u, t.wrap_v)
    n = {abi.TEX_R32F: 1, abi.TEX_RGB32F: 3}[t.format]
    assert (t.width, t.height) == (1, 1)
    return np.ctypeslib.as_array(C.cast(t.texels, abi.pf32), (n,)).copy(), (t.wrap_u, t.wrap_v)


def test_scene_file_matches_scene_reader_recipe(lib, resources):
    root, img, grid = resources
    sf = SceneFile(root / "scenes" / "scene.json", str(root))
    d = sf.desc()
    assert (d.n_materials, d.n_primitives, d.n_volumes, d.sort_and_group) == (5, 7, 1, 0)
    # ---- materials
    m = d.materials
    a, wrap = tex_value(d, m[0].albedo_tex)
    assert m[0].type == abi.MAT_MICROFACET and np.allclose(a, [.8, .8, .8]) and wrap == (abi.WRAP_CLAMP, abi.WRAP_CLAMP)
    assert tex_value(d, m[0].roughness_tex)[0] == np.float32(0.95) and tex_value(d, m[0].metallic_tex)[0] == 0 and m[0].normal_tex == -1
    a, wrap = tex_value(d, m[1].albedo_tex)
    assert np.array_equal(a, img) and wrap == (abi.WRAP_MIRROR, abi.WRAP_MIRROR)  # stbi RGBA8, mirror, top row first
    assert m[1].has_normal_flag == 1 and np.allclose(tex_value(d, m[1].normal_tex)[0], [.5, .5, 1])
    assert m[2].type == abi.MAT_EMITTER and list(m[2].li) == [10, 9, 8]
    assert m[3].type == abi.MAT_VOLUME and m[3].phase == abi.PHASE_HG and m[3].g == np.float32(0.3) and m[3].density_multiplier == 25
    assert list(m[3].scattering) == [np.float32(1.1)] * 3 and m[3].volume == 0
    v = d.volumes[0]
    assert (v.width, v.height, v.depth) == (6, 4, 5)
    assert np.array_equal(np.ctypeslib.as_array(v.dense, (5, 4, 6)), grid)
    # ---- primitives: same transforms / parameters as the SceneBuilder recipe (getTransform + inverse, Q25, sphere quirks)
    b = SceneBuilder()
    for n in ("white", "tex", "emitter", "cloud", "white2"):
        b.names[n] = {"white": 4, "tex": 1, "emitter": 2, "cloud": 3, "white2": 4}[n]  # the later "white" replaced the first
    b.add_volume("cloud", (0, 1.5, .5), (0, 20, 0), (2, 1.5, 2))
    b.add_rectangle("white", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    b.add_rectangle("tex", (0, 2, 3), (0, 0, 0), (8, 6, 1))
    b.add_mesh("tex", np.zeros((3, 3)), [[0, 1, 2]], None, (0, .3, 0), (10, 20, 30), (1, 2, 1))
    b.add_rectangle("emitter", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    b.add_sphere("emitter", (2.3, 2.5, 0), 0.05, collision=False)
    b.add_point("emitter", (0, 6, 0))
    for i, (p, q) in enumerate(zip(d.primitives[:7], b.primitives)):
        assert (p.type, p.material, p.collision) == (q.type, q.material, q.collision), i
        assert list(p.to_world) == list(q.to_world) and list(p.to_object) == list(q.to_object), i
        assert p.radius == q.radius and list(p.point) == list(q.point), i
    # ---- the OBJ: one vertex per face corner, fan triangulation, V flipped
    p = d.primitives[3]
    assert (p.n_triangles, p.n_vertices) == (3, 9)
    pos = np.ctypeslib.as_array(p.positions, (9, 3))
    uv = np.ctypeslib.as_array(p.uvs, (9, 2))
    idx = np.ctypeslib.as_array(p.indices, (9,))
    V = np.array([[-1, 0, -1], [1, 0, -1], [1, .5, 1], [-1, 0, 1]], np.float32)
    assert np.array_equal(idx, np.arange(9))
    assert np.array_equal(pos, V[[0, 1, 2, 0, 2, 3, 0, 1, 3]])
    assert np.array_equal(uv[:6], np.array([[0, 1], [1, 1], [1, 0], [0, 1], [1, 0], [0, 0]], np.float32)) and not uv[6:].any()
    # ---- camera and settings (Q26: up and aperture ignored, autoFocus = 3)
    st = sf.settings()
    assert (st.width, st.height, st.spp, st.bounces, st.hdr) == (96, 64, 8, 6, 0)
    cam = sf.camera()
    ref = scenes.CameraParams((0, 2, -5), (0, 1.2, 0), 40.0).make(96 / 64, lib)
    for f in ("position", "lower_left", "horizontal", "vertical", "side", "up"):
        assert list(getattr(cam, f)) == list(getattr(ref, f)), f
    assert cam.lens_radius == ref.lens_radius == np.float32(0.0001) / 2
    sf.close()


@pytest.mark.parametrize("mutate, needle", [
    (lambda s: s.pop("version"), "Version is not present"),
    (lambda s: s["materials"].append({"name": "d", "type": "diffuse"}), "invalid material type"),
    (lambda s: s["primitives"].append({"name": "b", "type": "box", "transform": {"position": [0, 0, 0], "scale": [1, 1, 1], "rotation": [0, 0, 0]}}), "invalid primitive type"),
    (lambda s: s["primitives"][1].__setitem__("materialName", "nope"), "materialName"),
    (lambda s: s["materials"][3].__setitem__("path", "vdb/cloud.vdb"), "OpenVDB"),
    (lambda s: s["camera"].pop("speed"), "camera.speed"),
])
def test_loader_errors_do_not_abort(lib, resources, mutate, needle):
    """Where SceneReader LOG(FATAL)s (SceneReader.cpp:24-41,221,647) the front end returns an error and a message."""
    root, _, _ = resources
    s = json.loads(json.dumps(SCENE))
    mutate(s)
    h = C.c_void_p()
    rc = lib.ne_b200_scene_file_parse(json.dumps(s).encode(), str(root).encode(), C.byref(h))
    assert rc == abi.ERR_INVALID and not h.value
    assert needle.lower() in lib.ne_b200_last_error().decode().lower()
    assert lib.ne_b200_scene_file_parse(b"{ not json", None, C.byref(h)) == abi.ERR_INVALID


def test_directional_light_material(lib):
    """SceneReader.cpp:156-168: le = albedo, direction = normalize(vec3(0) - position)."""
    s = {"version": "1", "materials": [{"name": "sun", "type": "directionalLight", "albedo": [1, 2, 3], "position": [3, 4, 0]}],
         "primitives": [{"name": "p", "type": "point", "materialName": "sun", "transform": {"position": [3, 4, 0], "scale": [1, 1, 1], "rotation": [0, 0, 0]}}],
         "camera": SCENE["camera"], "renderer": SCENE["renderer"]}
    sf = SceneFile(text=json.dumps(s), resources_dir="")
    m = sf.desc().materials[0]
    assert m.type == abi.MAT_DIRECTIONAL and list(m.li) == [1, 2, 3]
    assert np.allclose(list(m.direction), [-0.6, -0.8, 0], atol=1e-7)
    sf.close()


def test_infinite_area_light_material(lib, resources):
    """SceneReader.cpp:169-186: the map is loaded like any image texture (RGBA8, mirror)."""
    root, img, _ = resources
    s = {"version": "1", "materials": [{"name": "sky", "type": "infiniteAreaLight", "path": "imgs/check.png"}],
         "primitives": [{"name": "s", "type": "sphere", "radius": 30, "materialName": "sky", "transform": {"position": [0, 0, 0]}}],
         "camera": SCENE["camera"], "renderer": SCENE["renderer"]}
    sf = SceneFile(text=json.dumps(s), resources_dir=str(root))
    d = sf.desc()
    m = d.materials[0]
    assert m.type == abi.MAT_INFINITE and m.env_tex == 0
    a, wrap = tex_value(d, m.env_tex)
    assert np.array_equal(a, img) and wrap == (abi.WRAP_MIRROR, abi.WRAP_MIRROR)
    assert d.primitives[0].type == abi.PRIM_SPHERE and d.primitives[0].radius == 30
    sf.close()


def _gltf_doc(uri, nbytes):
    # two meshes under a small node tree (child before sibling: depth-first), the second without indices or uvs
    return {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 2]}],
            "nodes": [{"children": [1], "translation": [5, 5, 5]}, {"mesh": 0}, {"mesh": 1}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2}]},
                       {"primitives": [{"attributes": {"POSITION": 3}, "mode": 4}, {"attributes": {"POSITION": 3}, "mode": 1}]}],
            "buffers": [{"byteLength": nbytes, **({"uri": uri} if uri else {})}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 48}, {"buffer": 0, "byteOffset": 48, "byteLength": 32},
                            {"buffer": 0, "byteOffset": 80, "byteLength": 12}, {"buffer": 0, "byteOffset": 92, "byteLength": 36}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                          {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC2"},
                          {"bufferView": 2, "componentType": 5123, "count": 6, "type": "SCALAR"},
                          {"bufferView": 3, "componentType": 5126, "count": 3, "type": "VEC3"}]}


@pytest.mark.parametrize("container", ["base64", "bin", "glb"])
def test_gltf_reader(lib, tmp_path, container):
    """`gltf` primitives (SceneReader.cpp:245-267 -> assimp -> Model::processNode, Model.cpp:323-335): meshes of the node
    tree depth-first, no node transforms, indexed vertices kept, V flipped, non-triangle primitives skipped."""
    import base64
    P0 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    UV = np.array([[0, 0], [1, 0], [1, .25], [0, 1]], np.float32)
    I0 = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    P1 = np.array([[2, 0, 0], [3, 0, 1], [2, 1, 0]], np.float32)
    blob = P0.tobytes() + UV.tobytes() + I0.tobytes() + P1.tobytes()
    (tmp_path / "models").mkdir()
    if container == "base64":
        doc = _gltf_doc("data:application/octet-stream;base64," + base64.b64encode(blob).decode(), len(blob))
        (tmp_path / "models" / "m.gltf").write_text(json.dumps(doc))
        path = "models/m.gltf"
    elif container == "bin":
        (tmp_path / "models" / "m.bin").write_bytes(blob)
        (tmp_path / "models" / "m.gltf").write_text(json.dumps(_gltf_doc("m.bin", len(blob))))
        path = "models/m.gltf"
    else:
        js = json.dumps(_gltf_doc(None, len(blob))).encode()
        js += b" " * (-len(js) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
        (tmp_path / "models" / "m.glb").write_bytes(b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body)
        path = "models/m.glb"
    s = {"version": "1", "materials": [{"name": "m", "type": "microfacet", "roughness": 0.5, "metallic": 0.0, "albedo": [1, 1, 1]}],
         "primitives": [{"name": "g", "type": "gltf", "path": path, "materialName": "m",
                         "transform": {"position": [0, 0, 0], "scale": [1, 1, 1], "rotation": [0, 0, 0]}}],
         "camera": SCENE["camera"], "renderer": SCENE["renderer"]}
    sf = SceneFile(text=json.dumps(s), resources_dir=str(tmp_path))
    p = sf.desc().primitives[0]
    assert (p.type, p.n_vertices, p.n_triangles) == (abi.PRIM_MESH, 7, 3)
    assert np.array_equal(np.ctypeslib.as_array(p.positions, (7, 3)), np.concatenate([P0, P1]))
    assert np.array_equal(np.ctypeslib.as_array(p.indices, (9,)), [0, 1, 2, 0, 2, 3, 4, 5, 6])
    uv = np.ctypeslib.as_array(p.uvs, (7, 2))
    assert np.array_equal(uv[:4], np.stack([UV[:, 0], 1 - UV[:, 1]], 1)) and not uv[4:].any()
    sf.close()


def test_vol_reader_follows_the_reference_token_rules(lib, tmp_path):
    """ResourceManager.cpp:222-286: only space-terminated tokens count; line 2 is discarded; a value glued to a newline
    is dropped (std::stof stops at the newline); a last token without a trailing space is not captured."""
    p = tmp_path / "q.vol"
    p.write_text("2 2 1 \nthis line is discarded\n0.5 1.5\n2.5 3.5 4.5")
    g = read_vol(p, lib)
    assert g.shape == (1, 2, 2)
    assert np.array_equal(g.ravel(), np.array([0.5, 1.5, 3.5, 0], np.float32))  # 2.5 glued to 1.5's newline, 4.5 unterminated
    p.write_text("2 2 1\nx\n1 2 3 4 ")  # resolution without trailing spaces: the third number is never captured
    dims = (C.c_int32 * 3)()
    assert lib.ne_b200_vol_read(str(p).encode(), dims, None) == abi.ERR_INVALID
    rng = np.random.default_rng(1)
    grid = rng.random((3, 5, 7), dtype=np.float32) * 100
    write_vol(tmp_path / "rt.vol", grid, lib)
    assert np.array_equal(read_vol(tmp_path / "rt.vol", lib), grid)  # %.9g round-trips float32


def test_png_reader_matches_pillow(lib, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(2)
    for mode, shape in (("RGBA", (7, 5, 4)), ("RGB", (4, 9, 3)), ("L", (6, 6)), ("LA", (3, 8, 2)), ("P", (5, 5))):
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        im = Image.fromarray(a if mode != "P" else a % 16, mode)
        if mode == "P":
            im.putpalette(list(rng.integers(0, 256, 48, dtype=np.uint8)))
        im.save(tmp_path / f"{mode}.png")
        dims = (C.c_int32 * 2)()
        assert lib.ne_b200_image_read_png(str(tmp_path / f"{mode}.png").encode(), dims, None) == 0
        out = np.zeros((dims[1], dims[0], 4), np.uint8)
        assert lib.ne_b200_image_read_png(str(tmp_path / f"{mode}.png").encode(), dims, out.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        assert np.array_equal(out, np.asarray(im.convert("RGBA"))), mode  # what stbi_load(..., STBI_rgb_alpha) returns
    a16 = rng.integers(0, 65536, (4, 4), dtype=np.uint16)
    Image.fromarray(a16).save(tmp_path / "g16.png")
    dims = (C.c_int32 * 2)()
    out = np.zeros((4, 4, 4), np.uint8)
    assert lib.ne_b200_image_read_png(str(tmp_path / "g16.png").encode(), dims, out.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    assert np.array_equal(out[..., 0], (a16 >> 8).astype(np.uint8))


def test_image_writers(lib, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    img = (rng.random((9, 13, 3), dtype=np.float32) * 1.4 - 0.2).astype(np.float32)
    # PNG: clamp, truncating (uint8_t)(v * 255) (Texture.h:54-58)
    save_image(tmp_path / "o.png", img, lib)
    got = np.asarray(Image.open(tmp_path / "o.png"))
    want = (np.clip(img, 0, 1) * np.float32(255)).astype(np.uint8)
    assert got.shape == (9, 13, 3) and np.array_equal(got, want)
    # EXR: parse our own scanline file back (magic, channel list B G R float, uncompressed) and compare bit for bit
    save_image(tmp_path / "o.exr", img, lib)
    raw = (tmp_path / "o.exr").read_bytes()
    assert struct.unpack("<II", raw[:8]) == (20000630, 2)
    hdr_end = raw.index(b"screenWindowWidth\0float\0") + len(b"screenWindowWidth\0float\0") + 4 + 4 + 1
    assert b"channels\0chlist\0" in raw[:hdr_end] and raw[hdr_end - 1] == 0
    offs = struct.unpack(f"<{9}Q", raw[hdr_end:hdr_end + 72])
    out = np.zeros_like(img)
    for y, o in enumerate(offs):
        yy, size = struct.unpack("<iI", raw[o:o + 8])
        assert (yy, size) == (y, 13 * 12)
        planes = np.frombuffer(raw[o + 8:o + 8 + size], np.float32).reshape(3, 13)  # B, G, R
        out[y] = planes[::-1].T
    assert np.array_equal(out.view(np.uint32), img.view(np.uint32))
    assert offs[-1] + 8 + 13 * 12 == len(raw)
    # PPM: coreLoop's header, 16-bit big-endian, pixels last to first
    tm = np.clip(img, 0, 1)
    save_image(tmp_path / "o.ppm", tm, lib)
    raw = (tmp_path / "o.ppm").read_bytes()
    head = b"P6\n13 9\n65535\n"
    assert raw.startswith(head) and len(raw) == len(head) + 9 * 13 * 6
    px = np.frombuffer(raw[len(head):], ">u2").reshape(9 * 13, 3)
    assert np.array_equal(px, (tm.reshape(-1, 3)[::-1] * np.float32(65535)).astype(np.uint16))


# ---------------------------------------------------------------------------------------------------------------
# The scenes the reference ships (resources/scenes/**/*.json, 25 files). Runs where /root/reference exists.
# ---------------------------------------------------------------------------------------------------------------
REF_RESOURCES = "/root/reference/resources"


def shipped_scenes():
    import glob
    return sorted(glob.glob(os.path.join(REF_RESOURCES, "scenes", "**", "*.json"), recursive=True))


@pytest.mark.skipif(not os.path.isdir(REF_RESOURCES), reason="the reference tree is not present on this machine")
def test_every_shipped_scene_loads_or_fails_cleanly(lib):
    """Every JSON scene of the reference through ne_b200_scene_file_load: either a descriptor comes back, or an error code
    with a message naming what is missing (an absent .vol / .vdb / .obj asset, a material type the reference's own reader
    aborts on) - never an abort, never an exception across the C ABI (the reference LOG(FATAL)s)."""
    files = shipped_scenes()
    assert len(files) >= 20
    loaded, failed = [], {}
    for f in files:
        h = C.c_void_p()
        rc = lib.ne_b200_scene_file_load(f.encode(), (REF_RESOURCES + "/").encode(), C.byref(h))
        if rc == 0:
            d = lib.ne_b200_scene_file_desc(h).contents
            assert d.n_primitives >= 0 and d.n_materials >= 0  # empty.json is an empty scene
            st = abi.RenderSettings()
            assert lib.ne_b200_scene_file_settings(h, C.byref(st)) == 0 and st.width > 0 and st.height > 0
            lib.ne_b200_scene_file_free(h)
            loaded.append(os.path.basename(f))
        else:
            msg = lib.ne_b200_last_error().decode()
            assert rc in (abi.ERR_INVALID, abi.ERR_UNSUPPORTED, abi.ERR_NOMEM) and len(msg) > 8, (f, rc, msg)
            failed[os.path.basename(f)] = msg
    assert "surfaceAndLight.json" in loaded
    # the failures are about assets or features the reference tree itself lacks, not about the JSON
    for name, msg in failed.items():
        assert any(k in msg for k in ("couldn't read the file at", "needs OpenVDB", "invalid material type")), (name, msg)
    print(f"{len(loaded)} shipped scenes load ({', '.join(loaded)}); {len(failed)} fail cleanly")


@pytest.mark.skipif(not os.path.isdir(REF_RESOURCES), reason="the reference tree is not present on this machine")
def test_shipped_scene_fixture_is_current():
    """tests/golden/scene_surfaceAndLight.npz (made by tests/golden/make_scene_golden.py) holds the shipped scene's JSON
    text: it must still be the file the reference ships."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene_surfaceAndLight.npz"))
    assert bytes(g["scene_json"]).decode() == open(os.path.join(REF_RESOURCES, "scenes", "testing", "surfaceAndLight.json")).read()
