"""ctypes client of the ORACLE (oracle/_ref/libnarval_ref*.so = the reference's own TUs + oracle/ref/ref_harness.cpp).
Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
import ctypes as C
import os

import numpy as np

from narvalengine_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
f32, pf32, pu32, pi32 = C.c_float, abi.pf32, abi.pu32, abi.pi32


def _p(a, t=pf32):
    return a.ctypes.data_as(t)


def f32a(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float32)
    return a.reshape(shape) if shape else a


class RefOracle:
    def __init__(self, faithful=False):
        name = "libnarval_ref_faithful.so" if faithful else "libnarval_ref.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle ref` where /root/reference is present")
        L = self.lib = C.CDLL(path)
        L.neref_scene_create.restype = C.c_void_p
        L.neref_scene_create.argtypes = [C.POINTER(abi.SceneDesc)]
        L.neref_scene_destroy.argtypes = [C.c_void_p]
        L.neref_render.restype = C.c_double
        for n in ("neref_area_to_solid_angle", "neref_power_heuristic", "neref_roughness_to_alpha", "neref_ggx_D",
                  "neref_ggx_G", "neref_ggx_pdf", "neref_fresnel", "neref_hg_eval", "neref_isotropic_sample"):
            getattr(L, n).restype = C.c_float

    # ---- pure functions
    def tape(self, seed, n):
        out = np.zeros(n, np.float32)
        self.lib.neref_tape(C.c_uint32(seed), n, _p(out))
        return out

    def get_transform(self, pos, rot, scale):
        M, Mi = np.zeros(16, np.float32), np.zeros(16, np.float32)
        self.lib.neref_get_transform(_p(f32a(pos)), _p(f32a(rot)), _p(f32a(scale)), _p(M), _p(Mi))
        return M, Mi

    def transform_fn(self):
        return lambda p, r, s: tuple(list(map(float, m)) for m in self.get_transform(p, r, s))

    def onb(self, n):
        v, u = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.lib.neref_onb(_p(f32a(n)), _p(v), _p(u))
        return v, u

    def get_scale(self, M):
        s = np.zeros(3, np.float32)
        self.lib.neref_get_scale(_p(f32a(M)), _p(s))
        return s

    def area_to_solid_angle(self, pdf, n, p1, p2):
        return self.lib.neref_area_to_solid_angle(f32(pdf), _p(f32a(n)), _p(f32a(p1)), _p(f32a(p2)))

    def power_heuristic(self, a, b):
        return self.lib.neref_power_heuristic(f32(a), f32(b))

    def roughness_to_alpha(self, r):
        return self.lib.neref_roughness_to_alpha(f32(r))

    def sample_unit_sphere(self, e1, e2):
        o = np.zeros(3, np.float32)
        self.lib.neref_sample_unit_sphere(f32(e1), f32(e2), _p(o))
        return o

    def ggx_D(self, alpha, h):
        return self.lib.neref_ggx_D(f32(alpha), _p(f32a(h)))

    def ggx_G(self, alpha, wo, wi):
        return self.lib.neref_ggx_G(f32(alpha), _p(f32a(wo)), _p(f32a(wi)))

    def ggx_pdf(self, alpha, wi, h):
        return self.lib.neref_ggx_pdf(f32(alpha), _p(f32a(wi)), _p(f32a(h)))

    def fresnel(self, c):
        return self.lib.neref_fresnel(f32(c))

    def hg_eval(self, g, a, b):
        return self.lib.neref_hg_eval(f32(g), _p(f32a(a)), _p(f32a(b)))

    def hg_sample(self, g, seed):
        o = np.zeros(3, np.float32)
        self.lib.neref_hg_sample(f32(g), C.c_uint32(seed), _p(o))
        return o

    # ---- primitive-level entry points (the objects unitTests/tests.cpp exercises)
    def to_lcs(self, v, ns, ss, ts):
        o = np.zeros(3, np.float32)
        self.lib.neref_to_lcs(_p(f32a(v)), _p(f32a(ns)), _p(f32a(ss)), _p(f32a(ts)), _p(o))
        return o

    def to_world(self, v, ns, ss, ts):
        o = np.zeros(3, np.float32)
        self.lib.neref_to_world(_p(f32a(v)), _p(f32a(ns)), _p(f32a(ss)), _p(f32a(ts)), _p(o))
        return o

    def triangle_intersect(self, v0, v1, v2, o, d):
        t, p, n = np.zeros(2, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        did = self.lib.neref_triangle_intersect(_p(f32a(v0)), _p(f32a(v1)), _p(f32a(v2)), _p(f32a(o)), _p(f32a(d)), _p(t), _p(p), _p(n))
        return bool(did), t, p, n

    def triangle_barycentric(self, v0, v1, v2, p):
        o = np.zeros(3, np.float32)
        self.lib.neref_triangle_barycentric(_p(f32a(v0)), _p(f32a(v1)), _p(f32a(v2)), _p(f32a(p)), _p(o))
        return o

    def point_in_triangle_range(self, p, a, b, c):
        return bool(self.lib.neref_point_in_triangle_range(_p(f32a(p)), _p(f32a(a)), _p(f32a(b)), _p(f32a(c))))

    def aabb_intersect(self, bmin, bmax, o, d):
        t, p, n = np.zeros(2, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        did = self.lib.neref_aabb_intersect(_p(f32a(bmin)), _p(f32a(bmax)), _p(f32a(o)), _p(f32a(d)), _p(t), _p(p), _p(n))
        return bool(did), t, p, n

    def isotropic_sample(self, seed):
        o = np.zeros(3, np.float32)
        pdf = self.lib.neref_isotropic_sample(C.c_uint32(seed), _p(o))
        return o, pdf

    def tonemap(self, rgb):
        a = f32a(rgb).reshape(-1, 3)
        o = np.zeros_like(a)
        self.lib.neref_tonemap(_p(a), len(a), _p(o))
        return o

    def camera_make(self, cp, aspect):
        cam = abi.Camera()
        self.lib.neref_camera_make(_p(f32a(cp.look_from)), _p(f32a(cp.look_at)), _p(f32a(cp.up)), f32(cp.vfov),
                                   f32(aspect), f32(cp.aperture), f32(cp.focus), C.byref(cam))
        return cam

    def camera_rays(self, cp, aspect, seed, xy):
        xy = f32a(xy).reshape(-1, 2)
        n = len(xy)
        o, d = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        self.lib.neref_camera_rays(_p(f32a(cp.look_from)), _p(f32a(cp.look_at)), _p(f32a(cp.up)), f32(cp.vfov),
                                   f32(aspect), f32(cp.aperture), f32(cp.focus), C.c_uint32(seed), n, _p(xy), _p(o), _p(d))
        return o, d

    # ---- scenes
    def scene(self, builder_or_desc):
        desc = builder_or_desc.desc() if hasattr(builder_or_desc, "desc") else builder_or_desc
        h = self.lib.neref_scene_create(C.byref(desc))
        if not h:
            raise RuntimeError("neref_scene_create failed")
        return RefScene(self, h, builder_or_desc)


class FlatDesc:
    """Handle of a flattened scene; desc() is what Context.upload / RefOracle.scene take."""

    def __init__(self, ref_scene, handle):
        self.rs, self.h = ref_scene, C.c_void_p(handle)

    def desc(self):
        return self.rs.lib.neref_flat_desc(self.h).contents

    def close(self):
        if self.h:
            self.rs.lib.neref_flat_free(self.h)
            self.h = None


class RefScene:
    def __init__(self, oracle, handle, keep):
        self.o, self.h, self._keep = oracle, C.c_void_p(handle), keep
        self.lib = oracle.lib

    def close(self):
        self.lib.neref_scene_destroy(self.h)

    def flatten(self):
        """The reference-side adapter (oracle/ref/flatten.cpp) on this Scene*: a FlatDesc holding the ne_b200_scene_desc it
        produced (pointers alias this scene's buffers: keep the RefScene alive while it is used)."""
        self.lib.neref_flatten.restype = C.c_void_p
        self.lib.neref_flatten.argtypes = [C.c_void_p]
        self.lib.neref_flat_desc.restype = C.POINTER(abi.SceneDesc)
        self.lib.neref_flat_desc.argtypes = [C.c_void_p]
        self.lib.neref_flat_free.argtypes = [C.c_void_p]
        return FlatDesc(self, self.lib.neref_flatten(self.h))

    def counts(self):
        a, b = C.c_int(), C.c_int()
        self.lib.neref_scene_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def intersect(self, o, d, tmin=1e-11, tmax=float("inf")):
        o, d = f32a(o).reshape(-1, 3), f32a(d).reshape(-1, 3)
        hits = (abi.Hit * len(o))()
        self.lib.neref_intersect(self.h, len(o), _p(o), _p(d), f32(tmin), f32(tmax), hits)
        return hits

    def bsdf(self, instance, incoming, scattered, normals, uvs=None, seeds=None):
        i, s, nn = (f32a(x).reshape(-1, 3) for x in (incoming, scattered, normals))
        n = len(i)
        uv = f32a(uvs).reshape(-1, 2) if uvs is not None else np.zeros((n, 2), np.float32)
        ev, pdf, smp = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        sd = np.ascontiguousarray(seeds, np.uint32) if seeds is not None else None
        self.lib.neref_bsdf(self.h, n, instance, _p(i), _p(s), _p(nn), _p(uv), _p(sd, pu32) if sd is not None else None,
                            _p(ev), _p(pdf), _p(smp))
        return ev, pdf, smp

    def grid_tr(self, instance, o, d, tnear, tfar, seeds):
        o, d = f32a(o).reshape(-1, 3), f32a(d).reshape(-1, 3)
        n = len(o)
        tn, tf = f32a(tnear).reshape(n), f32a(tfar).reshape(n)
        sd = np.ascontiguousarray(seeds, np.uint32)
        tr, used = np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.neref_grid_tr(self.h, n, instance, _p(o), _p(d), _p(tn), _p(tf), _p(sd, pu32), _p(tr), _p(used, pi32))
        return tr, used

    def grid_sample(self, instance, o, d, tnear, tfar, seeds):
        o, d = f32a(o).reshape(-1, 3), f32a(d).reshape(-1, 3)
        n = len(o)
        tn, tf = f32a(tnear).reshape(n), f32a(tfar).reshape(n)
        sd = np.ascontiguousarray(seeds, np.uint32)
        T, so, sdir = (np.zeros((n, 3), np.float32) for _ in range(3))
        used = np.zeros(n, np.int32)
        self.lib.neref_grid_sample(self.h, n, instance, _p(o), _p(d), _p(tn), _p(tf), _p(sd, pu32), _p(T), _p(so),
                                   _p(sdir), _p(used, pi32))
        return T, so, sdir, used

    def density(self, instance, pts):
        p = f32a(pts).reshape(-1, 3)
        out, inv = np.zeros(len(p), np.float32), C.c_float()
        self.lib.neref_density(self.h, len(p), instance, _p(p), _p(out), C.byref(inv))
        return out, inv.value

    def li(self, o, d, bounces, seeds):
        o, d = f32a(o).reshape(-1, 3), f32a(d).reshape(-1, 3)
        n = len(o)
        sd = np.ascontiguousarray(seeds, np.uint32)
        L, used = np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
        self.lib.neref_li(self.h, n, _p(o), _p(d), bounces, _p(sd, pu32), _p(L), _p(used, pi32))
        return L, used

    def sample_one_light(self, incoming_dirs, hits, seeds):
        dd = f32a(incoming_dirs).reshape(-1, 3)
        n = len(dd)
        sd = np.ascontiguousarray(seeds, np.uint32)
        L, used = np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
        self.lib.neref_sample_one_light(self.h, n, _p(dd), hits, _p(sd, pu32), _p(L), _p(used, pi32))
        return L, used

    def render(self, cp, W, H, spp, bounces, seed=1, threads=1, rows=None, tonemapped=False, row_step=1):
        lin = np.zeros((H, W, 3), np.float32)
        tm = np.zeros((H, W, 3), np.float32) if tonemapped else None
        r0, r1 = rows if rows else (0, H)
        secs = self.lib.neref_render(self.h, _p(f32a(cp.look_from)), _p(f32a(cp.look_at)), _p(f32a(cp.up)),
                                     f32(cp.vfov), f32(cp.aperture), f32(cp.focus), W, H, spp, bounces,
                                     C.c_uint32(seed), threads, r0, r1, row_step, _p(lin), _p(tm) if tonemapped else None)
        return (lin, tm, secs) if tonemapped else (lin, secs)

    def render_tile(self, cp, W, H, spp, bounces, seed, tile):
        px = np.zeros((H, W, 3), np.float32)
        self.lib.neref_render_tile(self.h, _p(f32a(cp.look_from)), _p(f32a(cp.look_at)), _p(f32a(cp.up)), f32(cp.vfov),
                                   f32(cp.aperture), f32(cp.focus), W, H, spp, bounces, C.c_uint32(seed), tile, _p(px))
        return px
