"""Host-side mirror of the reference's renderer seam over the C ABI (include/ne_b200.h).

`Context` is a thin object wrapper of the `ne_b200_*` calls (one per GPU). `B200OfflineEngine` mirrors the public
surface of the reference's `OfflineEngine` (src/core/OfflineEngine.h:23-47): constructor `(camera, settings, scene)`,
`renderTile(index)` tile protocol over `numberOfTiles`, `pixels` (tone-mapped RGB32F, row-major W*y+x) — see
INTEGRATION.md for the C++ adapter a NarvalEngine maintainer would add.

There is no CPU implementation behind any of this: every call that computes goes to the CUDA library and raises
`NarvalB200Error` if it is missing or no GPU is visible.
"""
import ctypes as C

import numpy as np

from . import abi
from .abi import NarvalB200Error, check


def _p(a, t=abi.pf32):
    return a.ctypes.data_as(t)


def _f32(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


class SceneSettings:
    """SceneSettings of src/core/Settings.h:12-17."""

    def __init__(self, resolution=(512, 512), spp=16, bounces=6, hdr=False):
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.spp, self.bounces, self.hdr = int(spp), int(bounces), bool(hdr)


class Context:
    def __init__(self, device=0, lib=None, handle=None):
        """handle: wrap a context owned by someone else (a rank of a MultiContext); close() then leaves it alone."""
        self.lib = lib or abi.load_library()
        self._owned = handle is None
        if handle is None:
            h = C.c_void_p()
            check(self.lib, self.lib.ne_b200_create(device, C.byref(h)), "ne_b200_create")
            handle = h
        self.h = handle
        self._scene_keep = None

    def close(self):
        if self.h and self._owned:
            self.lib.ne_b200_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Run on the caller's CUDA stream (an int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        check(self.lib, self.lib.ne_b200_set_stream(self.h, C.c_void_p(cuda_stream)), "ne_b200_set_stream")

    # ---- scene / camera / render
    def upload(self, builder_or_desc):
        desc = builder_or_desc.desc() if hasattr(builder_or_desc, "desc") else builder_or_desc
        self._scene_keep = builder_or_desc
        check(self.lib, self.lib.ne_b200_scene_upload(self.h, C.byref(desc)), "ne_b200_scene_upload")

    def set_camera(self, cam):
        self._cam = cam
        check(self.lib, self.lib.ne_b200_camera_set(self.h, C.byref(cam)), "ne_b200_camera_set")

    def render(self, W, H, spp_begin, spp_end, bounces, seed=1, flags=0):
        check(self.lib, self.lib.ne_b200_render(self.h, W, H, spp_begin, spp_end, bounces, seed, flags), "ne_b200_render")

    def wait(self):
        check(self.lib, self.lib.ne_b200_wait(self.h), "ne_b200_wait")

    def clear(self):
        check(self.lib, self.lib.ne_b200_clear(self.h), "ne_b200_clear")

    def accum_buffer(self):
        p, n, s = C.c_void_p(), C.c_size_t(), C.c_int()
        check(self.lib, self.lib.ne_b200_accum_buffer(self.h, C.byref(p), C.byref(n), C.byref(s)), "ne_b200_accum_buffer")
        return p.value, n.value, s.value

    def set_samples_accumulated(self, n):
        check(self.lib, self.lib.ne_b200_set_samples_accumulated(self.h, n), "ne_b200_set_samples_accumulated")

    def accum_download(self, W, H):
        """Checkpoint: (sums[H,W,3], samples) of the accumulation buffer."""
        out, n = np.empty((H, W, 3), np.float32), C.c_int()
        check(self.lib, self.lib.ne_b200_accum_download(self.h, _p(out), C.byref(n)), "ne_b200_accum_download")
        return out, n.value

    def accum_upload(self, sums, samples):
        """Resume: continue rendering further sample ranges on top of a checkpoint."""
        a = np.ascontiguousarray(sums, np.float32)
        check(self.lib, self.lib.ne_b200_accum_upload(self.h, a.shape[1], a.shape[0], _p(a), samples), "ne_b200_accum_upload")

    def read_linear(self, W, H, out=None):
        out = np.empty((H, W, 3), np.float32) if out is None else out
        check(self.lib, self.lib.ne_b200_read_linear(self.h, _p(out)), "ne_b200_read_linear")
        return out

    def read_tonemapped(self, W, H, out=None):
        out = np.empty((H, W, 3), np.float32) if out is None else out
        check(self.lib, self.lib.ne_b200_read_tonemapped(self.h, _p(out)), "ne_b200_read_tonemapped")
        return out

    def render_frame(self, cam, W, H, spp, bounces, seed=1, flags=0, tonemapped=None, linear=None):
        """One whole frame through the single reference-facing call; buffers are HOST arrays (H,W,3) float32."""
        check(self.lib, self.lib.ne_b200_render_frame(self.h, C.byref(cam) if cam is not None else None, W, H, spp, bounces, seed,
                                                      flags, _p(tonemapped) if tonemapped is not None else None,
                                                      _p(linear) if linear is not None else None), "ne_b200_render_frame")

    def render_adaptive(self, cam, W, H, spp_min, spp_max, spp_batch, target_rel_mse, bounces, seed=1, flags=0, tonemapped=None, linear=None):
        """Progressive rendering with the stopping rule of ne_b200_render_adaptive; returns (spp_rendered, rel_mse_estimate, converged)."""
        r = abi.AdaptiveResult()
        check(self.lib, self.lib.ne_b200_render_adaptive(self.h, C.byref(cam) if cam is not None else None, W, H, spp_min, spp_max, spp_batch,
                                                         target_rel_mse, bounces, seed, flags, _p(tonemapped) if tonemapped is not None else None,
                                                         _p(linear) if linear is not None else None, C.byref(r)), "ne_b200_render_adaptive")
        return r.spp_rendered, r.rel_mse_estimate, bool(r.converged)

    def counters(self):
        c = abi.Counters()
        check(self.lib, self.lib.ne_b200_get_counters(self.h, C.byref(c)), "ne_b200_get_counters")
        return c

    def counters_reset(self):
        check(self.lib, self.lib.ne_b200_counters_reset(self.h), "ne_b200_counters_reset")

    # ---- test hooks
    def intersect(self, o, d, tmin=1e-11, tmax=float("inf")):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        hits = (abi.Hit * max(1, len(o)))()
        check(self.lib, self.lib.ne_b200_test_intersect(self.h, len(o), _p(o), _p(d), tmin, tmax, hits), "test_intersect")
        return hits

    def camera_rays(self, xy, tape):
        xy, tape = _f32(xy).reshape(-1, 2), _f32(tape).reshape(-1, 2)
        n = len(xy)
        o, d = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        check(self.lib, self.lib.ne_b200_test_camera_rays(self.h, n, _p(xy), _p(tape), _p(o), _p(d)), "test_camera_rays")
        return o, d

    def bsdf(self, instance, incoming, scattered, normals, uvs=None, tape=None):
        i, s, nn = (_f32(x).reshape(-1, 3) for x in (incoming, scattered, normals))
        n = len(i)
        uv = _f32(uvs).reshape(-1, 2) if uvs is not None else None
        tp = _f32(tape).reshape(-1, 2) if tape is not None else None
        ev, pdf, smp = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        check(self.lib, self.lib.ne_b200_test_bsdf(self.h, n, instance, _p(i), _p(s), _p(nn), _p(uv) if uv is not None else None,
                                                   _p(tp) if tp is not None else None, _p(ev), _p(pdf), _p(smp)), "test_bsdf")
        return ev, pdf, smp

    def grid_tr(self, instance, o, d, tnear, tfar, tape):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        n = len(o)
        tape = _f32(tape).reshape(n, -1)
        tr, used = np.zeros(n, np.float32), np.zeros(n, np.int32)
        check(self.lib, self.lib.ne_b200_test_grid_tr(self.h, n, instance, _p(o), _p(d), _p(_f32(tnear).reshape(n)), _p(_f32(tfar).reshape(n)),
                                                      _p(tape), tape.shape[1], _p(tr), _p(used, abi.pi32)), "test_grid_tr")
        return tr, used

    def grid_sample(self, instance, o, d, tnear, tfar, tape):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        n = len(o)
        tape = _f32(tape).reshape(n, -1)
        T, so, sd = (np.zeros((n, 3), np.float32) for _ in range(3))
        used = np.zeros(n, np.int32)
        check(self.lib, self.lib.ne_b200_test_grid_sample(self.h, n, instance, _p(o), _p(d), _p(_f32(tnear).reshape(n)),
                                                          _p(_f32(tfar).reshape(n)), _p(tape), tape.shape[1], _p(T), _p(so), _p(sd),
                                                          _p(used, abi.pi32)), "test_grid_sample")
        return T, so, sd, used

    def li_tape(self, o, d, bounces, tape):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        n = len(o)
        tape = _f32(tape).reshape(n, -1)
        L, used = np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
        check(self.lib, self.lib.ne_b200_test_li_tape(self.h, n, _p(o), _p(d), bounces, _p(tape), tape.shape[1], _p(L), _p(used, abi.pi32)),
              "test_li_tape")
        return L, used

    def li_philox(self, o, d, bounces, seed=1, flags=0):
        o, d = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3)
        L = np.zeros((len(o), 3), np.float32)
        check(self.lib, self.lib.ne_b200_test_li_philox(self.h, len(o), _p(o), _p(d), bounces, seed, flags, _p(L)), "test_li_philox")
        return L

    def sample_one_light(self, incoming_dirs, hits, tape):
        dd = _f32(incoming_dirs).reshape(-1, 3)
        n = len(dd)
        tape = _f32(tape).reshape(n, -1)
        L, used = np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
        check(self.lib, self.lib.ne_b200_test_sample_one_light(self.h, n, _p(dd), hits, _p(tape), tape.shape[1], _p(L), _p(used, abi.pi32)),
              "test_sample_one_light")
        return L, used

    def density(self, instance, pts):
        p = _f32(pts).reshape(-1, 3)
        out, inv = np.zeros(len(p), np.float32), C.c_float()
        check(self.lib, self.lib.ne_b200_test_density(self.h, len(p), instance, _p(p), _p(out), C.byref(inv)), "test_density")
        return out, inv.value

    def set_fast_shading(self, on):
        """The bsdf / sample_one_light / li_tape hooks run the production ("FAST") medium shading (csrc/ne_device.cuh)."""
        check(self.lib, self.lib.ne_b200_test_set_fast_shading(self.h, 1 if on else 0), "test_set_fast_shading")

    def philox(self, seed, pixel, sample, n):
        out = np.zeros(n, np.float32)
        check(self.lib, self.lib.ne_b200_test_philox(self.h, seed, pixel, sample, n, _p(out)), "test_philox")
        return out


class MultiContext:
    """Several GPUs of one box behind one handle in ONE process (ne_b200_create_multi, csrc/ne_multi.cu): scene replicas,
    sample-index partition, fused peer-memory reduce + resolve on the first device. `devices` may repeat an index."""

    def __init__(self, devices, lib=None):
        self.lib = lib or abi.load_library()
        ids = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(self.lib, self.lib.ne_b200_create_multi(ids, len(devices), C.byref(h)), "ne_b200_create_multi")
        self.h = h
        self._scene_keep = None

    def __len__(self):
        return self.lib.ne_b200_multi_count(self.h)

    def rank(self, r):
        return Context(lib=self.lib, handle=C.c_void_p(self.lib.ne_b200_multi_ctx(self.h, r)))

    def peer_access(self, r):
        return bool(self.lib.ne_b200_multi_peer_access(self.h, r))

    def upload(self, builder_or_desc):
        desc = builder_or_desc.desc() if hasattr(builder_or_desc, "desc") else builder_or_desc
        self._scene_keep = builder_or_desc
        check(self.lib, self.lib.ne_b200_multi_scene_upload(self.h, C.byref(desc)), "ne_b200_multi_scene_upload")

    def render(self, cam, W, H, spp, bounces, seed=1, flags=0):
        check(self.lib, self.lib.ne_b200_multi_render(self.h, C.byref(cam) if cam is not None else None, W, H, spp, bounces, seed, flags),
              "ne_b200_multi_render")

    def resolve(self, tonemapped=None, linear=None):
        check(self.lib, self.lib.ne_b200_multi_resolve(self.h, _p(tonemapped) if tonemapped is not None else None,
                                                       _p(linear) if linear is not None else None), "ne_b200_multi_resolve")

    def render_frame(self, cam, W, H, spp, bounces, seed=1, flags=0, tonemapped=None, linear=None):
        check(self.lib, self.lib.ne_b200_multi_render_frame(self.h, C.byref(cam) if cam is not None else None, W, H, spp, bounces, seed, flags,
                                                            _p(tonemapped) if tonemapped is not None else None,
                                                            _p(linear) if linear is not None else None), "ne_b200_multi_render_frame")

    def close(self):
        if self.h:
            self.lib.ne_b200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200OfflineEngine:
    """Drop-in for `OfflineEngine` (src/core/OfflineEngine.{h,cpp}) on one GPU, or on several (`device` = a list).

    Reference protocol (SceneEditor.cpp:547-604): the caller starts `renderTile(cam, index, finished)` for tile indices
    0..numberOfTiles.x*numberOfTiles.y-1 and re-uploads `pixels` whenever one finishes. Here the first `renderTile`
    call renders the WHOLE frame on the GPU; each call then publishes its tile's rows/columns of the resolved frame
    into `pixels` and returns True ("finished"). Unlike the reference (Q28) the last tile column/row also covers the
    W%40 / H%10 remainder, so `pixels` is fully defined."""

    def __init__(self, camera, settings, scene, device=0, seed=1, flags=0):
        self.ctx = MultiContext(device) if isinstance(device, (list, tuple)) else Context(device)
        self.seed, self.flags = seed, flags
        self.numberOfThreads = 16
        self.numberOfTiles = (40, 10)
        self.updateOfflineEngine(camera, settings, scene)

    def updateOfflineEngine(self, camera, settings, scene):
        self.camera, self.settings, self.scene = camera, settings, scene
        W, H = settings.resolution
        self.tileSize = (W // self.numberOfTiles[0], H // self.numberOfTiles[1])
        self.pixels = np.zeros((H, W, 3), np.float32)
        self.linear = np.zeros((H, W, 3), np.float32)
        self._frame = None
        self.ctx.upload(scene)
        cam = camera.make(W / H, self.ctx.lib) if hasattr(camera, "make") else camera
        self._cam = cam

    def _render_frame(self):
        W, H = self.settings.resolution
        tm = np.empty((H, W, 3), np.float32)
        self.ctx.render_frame(self._cam, W, H, self.settings.spp, self.settings.bounces, self.seed, self.flags, tm, self.linear)
        self._frame = tm

    def renderTile(self, index):
        if self._frame is None:
            self._render_frame()
        W, H = self.settings.resolution
        nx, ny = self.numberOfTiles
        mx, my = index % nx, index // nx
        x0, y0 = mx * self.tileSize[0], my * self.tileSize[1]
        x1 = W if mx == nx - 1 else x0 + self.tileSize[0]
        y1 = H if my == ny - 1 else y0 + self.tileSize[1]
        self.pixels[y0:y1, x0:x1] = self._frame[y0:y1, x0:x1]
        return True

    def render(self):
        """Whole frame: every tile in order."""
        for i in range(self.numberOfTiles[0] * self.numberOfTiles[1]):
            self.renderTile(i)
        return self.pixels

    def close(self):
        self.ctx.close()
