"""ctypes mirror of include/ne_b200.h (the C-ABI drop-in boundary) and loader of the CUDA library.

The product path has no CPU fallback: `load_library()` raises if `libnarval_b200.so` (built in-tree by
`__graft_entry__.build()` / `make -C narvalengine_b200/csrc`) is missing.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libnarval_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5

TEX_R32F, TEX_RG32F, TEX_RGB32F, TEX_RGBA32F, TEX_RGBA8 = range(5)
WRAP_CLAMP, WRAP_MIRROR = 0, 1
MAT_MICROFACET, MAT_EMITTER, MAT_VOLUME, MAT_DIRECTIONAL, MAT_INFINITE = range(5)
PHASE_ISOTROPIC, PHASE_HG = 0, 1
PRIM_RECTANGLE, PRIM_SPHERE, PRIM_POINT, PRIM_VOLUME, PRIM_MESH = range(5)
RENDER_GLOBAL_MAJORANT = 1
RENDER_MEGAKERNEL = 2

f32, i32, u32, u64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint64
pf32 = C.POINTER(C.c_float)
pi32 = C.POINTER(C.c_int32)
pu32 = C.POINTER(C.c_uint32)


class Texture(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("format", i32), ("wrap_u", i32), ("wrap_v", i32),
                ("texels", C.c_void_p)]


class Volume(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("depth", i32), ("dense", pf32), ("n_leaves", i32),
                ("leaf_origin", pi32), ("leaf_values", pf32)]


class Material(C.Structure):
    _fields_ = [("type", i32), ("albedo_tex", i32), ("roughness_tex", i32), ("metallic_tex", i32),
                ("normal_tex", i32), ("has_normal_flag", i32), ("li", f32 * 3), ("direction", f32 * 3),
                ("scattering", f32 * 3), ("absorption", f32 * 3), ("density_multiplier", f32), ("phase", i32),
                ("g", f32), ("volume", i32), ("env_tex", i32)]


class Primitive(C.Structure):
    _fields_ = [("type", i32), ("material", i32), ("to_world", f32 * 16), ("to_object", f32 * 16),
                ("radius", f32), ("point", f32 * 3), ("collision", i32), ("n_vertices", i32),
                ("n_triangles", i32), ("positions", pf32), ("uvs", pf32), ("indices", pu32)]


class SceneDesc(C.Structure):
    _fields_ = [("n_textures", i32), ("textures", C.POINTER(Texture)),
                ("n_volumes", i32), ("volumes", C.POINTER(Volume)),
                ("n_materials", i32), ("materials", C.POINTER(Material)),
                ("n_primitives", i32), ("primitives", C.POINTER(Primitive)),
                ("sort_and_group", i32)]


class Camera(C.Structure):
    _fields_ = [("position", f32 * 3), ("lower_left", f32 * 3), ("horizontal", f32 * 3), ("vertical", f32 * 3),
                ("side", f32 * 3), ("up", f32 * 3), ("lens_radius", f32)]


class RenderSettings(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("spp", i32), ("bounces", i32), ("hdr", i32)]


class Hit(C.Structure):
    _fields_ = [("hit_point", f32 * 3), ("normal", f32 * 3), ("uv", f32 * 2), ("t_near", f32), ("t_far", f32),
                ("hit", i32), ("instance", i32), ("is_light", i32), ("primitive", i32)]


class Counters(C.Structure):
    _fields_ = [(n, u64) for n in ("paths", "extend_rays", "shadow_rays", "delta_steps", "ratio_steps",
                                   "brick_visits", "bvh_nodes", "tri_tests", "prim_tests", "scatter_events",
                                   "surface_events", "wavefront_iterations", "kernel_launches")] + \
               [(n, C.c_double) for n in ("ms_render", "ms_volume_kernel", "ms_extend_kernel", "ms_shade_kernel",
                                          "ms_upload")] + \
               [(n, u32) for n in ("bytes_per_tracking_step", "bytes_per_bvh_node", "bytes_per_triangle",
                                   "bytes_per_path_record")] + \
               [("ms_other_kernel", C.c_double)]


class AdaptiveResult(C.Structure):
    _fields_ = [("spp_rendered", i32), ("rel_mse_estimate", f32), ("converged", i32)]


# Every symbol include/ne_b200.h declares: name -> (restype, argtypes). tests/test_abi.py checks the list
# against the header and that the built library exports each one.
_ctx = C.c_void_p
SYMBOLS = {
    "ne_b200_make_transform": (C.c_int, [pf32, pf32, pf32, pf32, pf32]),
    "ne_b200_camera_make": (C.c_int, [pf32, pf32, pf32, f32, f32, f32, f32, C.POINTER(Camera)]),
    "ne_b200_host_build_bricks": (C.c_int, [C.POINTER(Volume), pi32, pi32, pf32, pf32, pf32]),
    "ne_b200_host_build_bvh": (C.c_int, [pf32, i32, pu32, i32, pi32, C.c_void_p, pf32]),
    "ne_b200_scene_file_load": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "ne_b200_scene_file_parse": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "ne_b200_scene_file_desc": (C.POINTER(SceneDesc), [C.c_void_p]),
    "ne_b200_scene_file_camera": (C.c_int, [C.c_void_p, C.POINTER(Camera)]),
    "ne_b200_scene_file_settings": (C.c_int, [C.c_void_p, C.POINTER(RenderSettings)]),
    "ne_b200_scene_file_free": (None, [C.c_void_p]),
    "ne_b200_vol_read": (C.c_int, [C.c_char_p, pi32, pf32]),
    "ne_b200_vol_write": (C.c_int, [C.c_char_p, pi32, pf32]),
    "ne_b200_image_read_png": (C.c_int, [C.c_char_p, pi32, C.POINTER(C.c_uint8)]),
    "ne_b200_image_write_png": (C.c_int, [C.c_char_p, C.c_int, C.c_int, pf32]),
    "ne_b200_image_write_exr": (C.c_int, [C.c_char_p, C.c_int, C.c_int, pf32]),
    "ne_b200_image_write_ppm": (C.c_int, [C.c_char_p, C.c_int, C.c_int, pf32]),
    "ne_b200_last_error": (C.c_char_p, []),
    "ne_b200_device_count": (C.c_int, []),
    "ne_b200_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "ne_b200_destroy": (None, [_ctx]),
    "ne_b200_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "ne_b200_scene_upload": (C.c_int, [_ctx, C.POINTER(SceneDesc)]),
    "ne_b200_camera_set": (C.c_int, [_ctx, C.POINTER(Camera)]),
    "ne_b200_render": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u64, u32]),
    "ne_b200_wait": (C.c_int, [_ctx]),
    "ne_b200_clear": (C.c_int, [_ctx]),
    "ne_b200_accum_buffer": (C.c_int, [_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "ne_b200_set_samples_accumulated": (C.c_int, [_ctx, C.c_int]),
    "ne_b200_accum_download": (C.c_int, [_ctx, pf32, C.POINTER(C.c_int)]),
    "ne_b200_accum_upload": (C.c_int, [_ctx, C.c_int, C.c_int, pf32, C.c_int]),
    "ne_b200_read_linear": (C.c_int, [_ctx, pf32]),
    "ne_b200_read_tonemapped": (C.c_int, [_ctx, pf32]),
    "ne_b200_render_frame": (C.c_int, [_ctx, C.POINTER(Camera), C.c_int, C.c_int, C.c_int, C.c_int, u64, u32,
                                       pf32, pf32]),
    "ne_b200_create_multi": (C.c_int, [pi32, C.c_int, C.POINTER(C.c_void_p)]),
    "ne_b200_multi_destroy": (None, [C.c_void_p]),
    "ne_b200_multi_count": (C.c_int, [C.c_void_p]),
    "ne_b200_multi_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ne_b200_multi_peer_access": (C.c_int, [C.c_void_p, C.c_int]),
    "ne_b200_multi_scene_upload": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc)]),
    "ne_b200_multi_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_int, C.c_int, u64, u32]),
    "ne_b200_multi_resolve": (C.c_int, [C.c_void_p, pf32, pf32]),
    "ne_b200_multi_render_frame": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_int, C.c_int, u64, u32,
                                             pf32, pf32]),
    "ne_b200_render_adaptive": (C.c_int, [_ctx, C.POINTER(Camera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32, C.c_int, u64, u32,
                                          pf32, pf32, C.POINTER(AdaptiveResult)]),
    "ne_b200_get_counters": (C.c_int, [_ctx, C.POINTER(Counters)]),
    "ne_b200_counters_reset": (C.c_int, [_ctx]),
    "ne_b200_test_intersect": (C.c_int, [_ctx, C.c_int, pf32, pf32, f32, f32, C.POINTER(Hit)]),
    "ne_b200_test_camera_rays": (C.c_int, [_ctx, C.c_int, pf32, pf32, pf32, pf32]),
    "ne_b200_test_bsdf": (C.c_int, [_ctx, C.c_int, C.c_int, pf32, pf32, pf32, pf32, pf32, pf32, pf32, pf32]),
    "ne_b200_test_grid_tr": (C.c_int, [_ctx, C.c_int, C.c_int, pf32, pf32, pf32, pf32, pf32, C.c_int, pf32, pi32]),
    "ne_b200_test_grid_sample": (C.c_int, [_ctx, C.c_int, C.c_int, pf32, pf32, pf32, pf32, pf32, C.c_int, pf32,
                                           pf32, pf32, pi32]),
    "ne_b200_test_li_tape": (C.c_int, [_ctx, C.c_int, pf32, pf32, C.c_int, pf32, C.c_int, pf32, pi32]),
    "ne_b200_test_li_philox": (C.c_int, [_ctx, C.c_int, pf32, pf32, C.c_int, u64, u32, pf32]),
    "ne_b200_test_sample_one_light": (C.c_int, [_ctx, C.c_int, pf32, C.POINTER(Hit), pf32, C.c_int, pf32, pi32]),
    "ne_b200_test_density": (C.c_int, [_ctx, C.c_int, C.c_int, pf32, pf32, pf32]),
    "ne_b200_test_read_bricks": (C.c_int, [_ctx, C.c_int, pi32, pi32, pf32, pf32, pf32]),
    "ne_b200_test_set_fast_shading": (C.c_int, [_ctx, C.c_int]),
    "ne_b200_test_philox": (C.c_int, [_ctx, u64, u32, u32, C.c_int, pf32]),
}

_lib = None


class NarvalB200Error(RuntimeError):
    pass


def load_library(path=None):
    """Load libnarval_b200.so and bind every declared symbol. Raises if the library is missing: there is no
    CPU implementation of the render path to fall back to."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NarvalB200Error(
            f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "narvalengine_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib, rc, what=""):
    if rc != OK:
        msg = lib.ne_b200_last_error()
        raise NarvalB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
