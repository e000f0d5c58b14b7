"""Host-side scene description builder: the Python mirror of what SceneReader::processMaterial /
processPrimitives / processCameraAndRenderer assemble (reference src/io/SceneReader.cpp:67-675), producing the
POD `ne_b200_scene_desc` of include/ne_b200.h. Numpy arrays referenced by the descriptor are kept alive by the
builder object.
"""
import ctypes as C
import numpy as np

from . import abi


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class SceneBuilder:
    """Accumulates textures, volumes, materials and primitives in JSON order.

    `transform_fn(pos, rot_deg, scale) -> (to_world[16], to_object[16])` defaults to the library's
    ne_b200_make_transform (getTransform + glm::inverse, reference Math.h:852-859)."""

    def __init__(self, transform_fn=None, sort_and_group=False):
        self.textures, self.volumes, self.materials, self.primitives = [], [], [], []
        self.names = {}
        self._keep = []
        self._transform_fn = transform_fn
        self.sort_and_group = sort_and_group

    # -- helpers ---------------------------------------------------------------------------------------------
    def _transform(self, pos, rot, scale):
        fn = self._transform_fn
        if fn is None:
            lib = abi.load_library()

            def fn(p, r, s):
                M = (C.c_float * 16)()
                Mi = (C.c_float * 16)()
                abi.check(lib, lib.ne_b200_make_transform(_f3(p), _f3(r), _f3(s), M, Mi), "make_transform")
                return list(M), list(Mi)
        return fn(pos, rot, scale)

    def _arr(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        self._keep.append(a)
        return a

    # -- textures / volumes ----------------------------------------------------------------------------------
    def add_texture(self, data, fmt, width=1, height=1, wrap=abi.WRAP_CLAMP):
        dtype = np.uint8 if fmt == abi.TEX_RGBA8 else np.float32
        a = self._arr(data, dtype)
        t = abi.Texture(width, height, fmt, wrap, wrap, a.ctypes.data_as(C.c_void_p))
        self.textures.append(t)
        return len(self.textures) - 1

    def add_volume_dense(self, grid):
        """grid: numpy array indexed [z, y, x] (x fastest), the Texture(W,H,D,R32F) of loadVolasTexture."""
        g = self._arr(grid, np.float32)
        d, h, w = g.shape
        self.volumes.append(abi.Volume(w, h, d, g.ctypes.data_as(abi.pf32), 0, None, None))
        return len(self.volumes) - 1

    def add_volume_leaves(self, dims, origins, values):
        """dims=(W,H,D); origins: [n,3] int32 (multiples of 8, texture space); values: [n,8,8,8] indexed [z,y,x]."""
        o = self._arr(origins, np.int32)
        v = self._arr(values, np.float32)
        self.volumes.append(abi.Volume(dims[0], dims[1], dims[2], None, len(o), o.ctypes.data_as(abi.pi32),
                                       v.ctypes.data_as(abi.pf32)))
        return len(self.volumes) - 1

    # -- materials (SceneReader.cpp:67-222) ------------------------------------------------------------------
    def _mat(self, name, m):
        self.materials.append(m)
        self.names[name] = len(self.materials) - 1
        return self.names[name]

    def add_microfacet(self, name, albedo, roughness, metallic, normal_map=None, albedo_image=None, normal_image=None):
        """albedo: (r,g,b) -> 1x1 RGB32F clamp texture, or albedo_image=(rgba8 HxWx4 array) -> RGBA8 mirror.
        normal_map: (x,y,z) -> the 1x1 RGB32F texture SceneReader builds for "normalMap" (SceneReader.cpp:117-121);
        normal_image: an RGBA8 image attached as TextureName::NORMAL (mirror wrap, like any loaded image)."""
        m = abi.Material()
        m.type = abi.MAT_MICROFACET
        if albedo_image is not None:
            img = np.asarray(albedo_image, dtype=np.uint8)
            m.albedo_tex = self.add_texture(img, abi.TEX_RGBA8, img.shape[1], img.shape[0], abi.WRAP_MIRROR)
        else:
            m.albedo_tex = self.add_texture(albedo, abi.TEX_RGB32F)
        m.metallic_tex = self.add_texture([metallic], abi.TEX_R32F)
        m.roughness_tex = self.add_texture([roughness], abi.TEX_R32F)
        m.normal_tex = -1
        m.has_normal_flag = 0
        if normal_map is not None:
            m.normal_tex = self.add_texture(normal_map, abi.TEX_RGB32F)
            m.has_normal_flag = 1  # NORMAL is the last addTexture call (Q24)
        if normal_image is not None:
            img = np.asarray(normal_image, dtype=np.uint8)
            m.normal_tex = self.add_texture(img, abi.TEX_RGBA8, img.shape[1], img.shape[0], abi.WRAP_MIRROR)
            m.has_normal_flag = 1
        m.volume = -1
        m.env_tex = -1
        return self._mat(name, m)

    def add_emitter(self, name, li):
        m = abi.Material()
        m.type = abi.MAT_EMITTER
        m.li = _f3(li)
        m.albedo_tex = m.roughness_tex = m.metallic_tex = m.normal_tex = -1
        m.volume = -1
        m.env_tex = -1
        return self._mat(name, m)

    def add_infinite_area_light(self, name, env_image):
        """infiniteAreaLight material (SceneReader.cpp:169-186): a lat-long RGBA8 map loaded like any image texture
        (mirror wrap, nearest texel)."""
        img = np.asarray(env_image, dtype=np.uint8)
        m = abi.Material()
        m.type = abi.MAT_INFINITE
        m.env_tex = self.add_texture(img, abi.TEX_RGBA8, img.shape[1], img.shape[0], abi.WRAP_MIRROR)
        m.albedo_tex = m.roughness_tex = m.metallic_tex = m.normal_tex = -1
        m.volume = -1
        return self._mat(name, m)

    def add_directional(self, name, le, position):
        """directionalLight material (SceneReader.cpp:156-168): le = albedo, direction = normalize(-position)."""
        m = abi.Material()
        m.type = abi.MAT_DIRECTIONAL
        m.li = _f3(le)
        p = np.asarray(position, np.float32)
        m.direction = _f3((np.float32(0) - p) / np.sqrt((p * p).sum(dtype=np.float32)))
        m.albedo_tex = m.roughness_tex = m.metallic_tex = m.normal_tex = -1
        m.volume = -1
        m.env_tex = -1
        return self._mat(name, m)

    def add_volume_material(self, name, scattering, absorption, density, volume, phase="isotropic", g=0.0):
        m = abi.Material()
        m.type = abi.MAT_VOLUME
        m.scattering = _f3(scattering)
        m.absorption = _f3(absorption)
        m.density_multiplier = density
        m.phase = abi.PHASE_HG if phase in ("hg", "henyey-greenstein") else abi.PHASE_ISOTROPIC
        m.g = g
        m.volume = volume
        m.albedo_tex = m.roughness_tex = m.metallic_tex = m.normal_tex = -1
        m.env_tex = -1
        return self._mat(name, m)

    # -- primitives (SceneReader.cpp:224-648) ----------------------------------------------------------------
    def _prim(self, ptype, material, pos, rot, scale):
        p = abi.Primitive()
        p.type = ptype
        p.material = self.names[material] if isinstance(material, str) else material
        M, Mi = self._transform(pos, rot, scale)
        p.to_world = (C.c_float * 16)(*M)
        p.to_object = (C.c_float * 16)(*Mi)
        p.collision = 1
        self.primitives.append(p)
        return p

    def add_rectangle(self, material, pos, rot=(0, 0, 0), scale=(1, 1, 1)):
        return self._prim(abi.PRIM_RECTANGLE, material, pos, rot, scale)

    def add_sphere(self, material, pos, radius, collision=True):
        p = self._prim(abi.PRIM_SPHERE, material, pos, (0, 0, 0), (1, 1, 1))  # scale/rotation ignored (:378)
        p.radius = radius
        p.collision = 1 if collision else 0
        return p

    def add_point(self, material, pos, rot=(0, 0, 0), scale=(1, 1, 1)):
        p = self._prim(abi.PRIM_POINT, material, pos, rot, scale)
        p.point = _f3(pos)  # Q25: vertex = pos and the transform translates by pos again
        return p

    def add_volume(self, material, pos, rot=(0, 0, 0), scale=(1, 1, 1)):
        return self._prim(abi.PRIM_VOLUME, material, pos, rot, scale)

    def add_mesh(self, material, positions, indices, uvs=None, pos=(0, 0, 0), rot=(0, 0, 0), scale=(1, 1, 1)):
        p = self._prim(abi.PRIM_MESH, material if material is not None else -1, pos, rot, scale)
        v = self._arr(np.asarray(positions).reshape(-1, 3), np.float32)
        i = self._arr(np.asarray(indices).reshape(-1, 3), np.uint32)
        p.n_vertices, p.n_triangles = len(v), len(i)
        p.positions = v.ctypes.data_as(abi.pf32)
        p.indices = i.ctypes.data_as(abi.pu32)
        if uvs is not None:
            u = self._arr(np.asarray(uvs).reshape(-1, 2), np.float32)
            p.uvs = u.ctypes.data_as(abi.pf32)
        return p

    # -- descriptor ------------------------------------------------------------------------------------------
    def desc(self):
        d = abi.SceneDesc()

        def arr(items, typ):
            a = (typ * max(1, len(items)))(*items)
            self._keep.append(a)
            return a
        d.n_textures = len(self.textures)
        d.textures = arr(self.textures, abi.Texture)
        d.n_volumes = len(self.volumes)
        d.volumes = arr(self.volumes, abi.Volume)
        d.n_materials = len(self.materials)
        d.materials = arr(self.materials, abi.Material)
        d.n_primitives = len(self.primitives)
        d.primitives = arr(self.primitives, abi.Primitive)
        d.sort_and_group = 1 if self.sort_and_group else 0
        self._keep.append(d)
        return d


class CameraParams:
    """Arguments of Camera::Camera as SceneReader::processCameraAndRenderer passes them (SceneReader.cpp:650-668):
    up=(0,1,0), aperture=1e-4 and autoFocus -> focus 3 (Q26)."""

    def __init__(self, look_from, look_at, vfov, up=(0, 1, 0), aperture=1e-4, focus=3.0):
        self.look_from, self.look_at, self.up = tuple(look_from), tuple(look_at), tuple(up)
        self.vfov, self.aperture, self.focus = float(vfov), float(aperture), float(focus)

    def make(self, aspect, lib=None):
        lib = lib or abi.load_library()
        cam = abi.Camera()
        abi.check(lib, lib.ne_b200_camera_make(_f3(self.look_from), _f3(self.look_at), _f3(self.up), self.vfov,
                                               aspect, self.aperture, self.focus, C.byref(cam)), "camera_make")
        return cam


class SceneFile:
    """A JSON scene file loaded by the library's front end (ne_b200_scene_file_*, csrc/ne_frontend.cpp): the
    counterpart of `SceneReader(filePath)` + `getScene() / getMainCamera() / getSettings()` (reference
    src/io/SceneReader.cpp:10-53,676-690). Can be handed to Context.upload() like a SceneBuilder."""

    def __init__(self, path=None, resources_dir="", text=None, lib=None):
        self.lib = lib or abi.load_library()
        h = C.c_void_p()
        res = resources_dir.encode() if resources_dir is not None else None
        if text is not None:
            rc = self.lib.ne_b200_scene_file_parse(text.encode(), res, C.byref(h))
        else:
            rc = self.lib.ne_b200_scene_file_load(str(path).encode(), res, C.byref(h))
        abi.check(self.lib, rc, "ne_b200_scene_file_load")
        self.h = h

    def desc(self):
        return self.lib.ne_b200_scene_file_desc(self.h).contents

    def camera(self):
        cam = abi.Camera()
        abi.check(self.lib, self.lib.ne_b200_scene_file_camera(self.h, C.byref(cam)), "ne_b200_scene_file_camera")
        return cam

    def settings(self):
        st = abi.RenderSettings()
        abi.check(self.lib, self.lib.ne_b200_scene_file_settings(self.h, C.byref(st)), "ne_b200_scene_file_settings")
        return st

    def close(self):
        if self.h:
            self.lib.ne_b200_scene_file_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_vol(path, lib=None):
    """ResourceManager::loadVolasTexture's .vol parser -> array indexed [z, y, x]."""
    lib = lib or abi.load_library()
    dims = (C.c_int32 * 3)()
    abi.check(lib, lib.ne_b200_vol_read(str(path).encode(), dims, None), "ne_b200_vol_read")
    g = np.zeros((dims[2], dims[1], dims[0]), np.float32)
    abi.check(lib, lib.ne_b200_vol_read(str(path).encode(), dims, g.ctypes.data_as(abi.pf32)), "ne_b200_vol_read")
    return g


def write_vol(path, grid, lib=None):
    lib = lib or abi.load_library()
    g = np.ascontiguousarray(grid, np.float32)
    dims = (C.c_int32 * 3)(g.shape[2], g.shape[1], g.shape[0])
    abi.check(lib, lib.ne_b200_vol_write(str(path).encode(), dims, g.ctypes.data_as(abi.pf32)), "ne_b200_vol_write")


def save_image(path, rgb, lib=None):
    """saveImage(pixels, W, H, RGB32F, PNG|EXR) (reference materials/Texture.h:44-75) / output.ppm by extension."""
    lib = lib or abi.load_library()
    a = np.ascontiguousarray(rgb, np.float32)
    h, w = a.shape[:2]
    fn = {".png": lib.ne_b200_image_write_png, ".exr": lib.ne_b200_image_write_exr, ".ppm": lib.ne_b200_image_write_ppm}[
        str(path)[-4:].lower()]
    abi.check(lib, fn(str(path).encode(), w, h, a.ctypes.data_as(abi.pf32)), "ne_b200_image_write")
