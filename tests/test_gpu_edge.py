"""Degenerate and ragged inputs through the production path on the GPU: nothing may hang, crash or produce non-finite
radiance, and state errors are reported, not asserted (the reference LOG(FATAL)s / throws in several of these)."""
import ctypes as C

import numpy as np
import pytest

import scenes
from narvalengine_b200 import abi
from narvalengine_b200.engine import Context
from narvalengine_b200.scene import SceneBuilder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def frame(ctx, b, cam, W, H, spp, flags=0):
    ctx.upload(b)
    lin = np.full((H, W, 3), -1, np.float32)
    ctx.render_frame(cam.make(W / H, ctx.lib), W, H, spp, 6, 1, flags, None, lin)
    return lin


def test_call_order_errors():
    c = Context(0)
    assert c.lib.ne_b200_render(c.h, 8, 8, 0, 1, 6, 1, 0) == abi.ERR_STATE            # no scene
    out = np.zeros((8, 8, 3), np.float32)
    assert c.lib.ne_b200_read_linear(c.h, out.ctypes.data_as(abi.pf32)) == abi.ERR_STATE  # nothing rendered
    c.upload(scenes.cornell_c1())
    assert c.lib.ne_b200_render(c.h, 8, 8, 0, 1, 6, 1, 0) == abi.ERR_STATE            # no camera
    c.set_camera(scenes.CORNELL_CAMERA.make(1.0, c.lib))
    assert c.lib.ne_b200_render(c.h, 0, 8, 0, 1, 6, 1, 0) == abi.ERR_INVALID
    assert c.lib.ne_b200_render(c.h, 8, 8, 3, 1, 6, 1, 0) == abi.ERR_INVALID
    assert c.lib.ne_b200_render(c.h, 8, 8, 0, 0, 6, 1, 0) == abi.OK                   # empty sample range: allocates, renders nothing
    assert c.lib.ne_b200_read_linear(c.h, out.ctypes.data_as(abi.pf32)) == abi.ERR_STATE
    c.close()


def test_empty_scene_and_scene_without_lights(ctx):
    cam = scenes.CORNELL_CAMERA
    assert not frame(ctx, SceneBuilder(), cam, 16, 8, 4).any()                        # nothing to hit: black, no hang
    b = SceneBuilder()
    b.add_microfacet("floor", (.8, .8, .8), 0.9, 0.0)
    b.add_rectangle("floor", (0, 0, 0), (90, 0, 0), (8, 8, 1))
    img = frame(ctx, b, cam, 16, 8, 4)                                                # Scene::lights empty: the reference throws std::out_of_range
    assert np.isfinite(img).all() and not img.any()
    b = SceneBuilder()
    b.add_emitter("light", (5, 6, 7))
    b.add_rectangle("light", (0, 2, 0), (0, 0, 0), (4, 4, 1))
    img = frame(ctx, b, cam, 16, 8, 4)                                                # only an emitter: camera rays see Li exactly
    cover = img / np.array([5, 6, 7], np.float32)                                     # = (samples that hit the light) / spp
    assert np.isfinite(img).all() and cover.max() == 1 and set(np.unique(cover)) <= {0.0, 0.25, 0.5, 0.75, 1.0}
    assert np.array_equal(cover[..., 0], cover[..., 1]) and np.array_equal(cover[..., 0], cover[..., 2])


@pytest.mark.parametrize("flags", [0, abi.RENDER_MEGAKERNEL])
def test_all_zero_and_tiny_volumes(ctx, flags):
    cam = scenes.CameraParams((0, 1, -6), (0, 1, 0), 45.0)
    for grid in (np.zeros((5, 3, 9), np.float32), np.ones((1, 1, 1), np.float32), np.zeros((8, 8, 8), np.float32)):
        b = scenes.noise_volume_scene(res=grid.shape[::-1], density=10.0, light="rect", grid=grid)
        img = frame(ctx, b, cam, 24, 16, 8, flags)
        assert np.isfinite(img).all() and (img >= 0).all()
    o = np.zeros((0, 3), np.int32)
    v = np.zeros((0, 8, 8, 8), np.float32)
    b = SceneBuilder()
    vol = b.add_volume_leaves((16, 16, 16), o, v)                                     # a VDB with no active leaf
    b.add_volume_material("cloud", (1.1, 1.1, 1.1), (.01, .01, .01), 10.0, vol)
    b.add_emitter("light", (50, 50, 50))
    b.add_volume("cloud", (0, 1, 0), (0, 0, 0), (2, 2, 2))
    b.add_rectangle("light", (0, 3.9, 0), (-89, 0, 0), (1, 1, 1))
    assert np.isfinite(frame(ctx, b, cam, 24, 16, 8, flags)).all()


def test_ragged_frames_and_sample_ranges(ctx):
    b = scenes.mixed_scene()
    cam = scenes.MIXED_CAMERA
    one = frame(ctx, b, cam, 1, 1, 64)
    assert one.shape == (1, 1, 3) and np.isfinite(one).all()
    odd = frame(ctx, b, cam, 37, 23, 5)
    assert np.isfinite(odd).all() and odd.mean() > 0
    # far-away sample indices are as good as the first ones (Philox counter = sample index)
    ctx.set_camera(cam.make(37 / 23, ctx.lib))
    ctx.clear()
    ctx.render(37, 23, 2_000_000_000, 2_000_000_005, 6, seed=7)
    ctx.wait()
    far = ctx.read_linear(37, 23)
    assert np.isfinite(far).all() and abs(far.mean() - odd.mean()) / odd.mean() < 0.5


def test_reupload_growing_and_shrinking_scenes(ctx):
    cam = scenes.CameraParams((0, 1, -6), (0, 1, 0), 45.0)
    means = []
    for res in ((16, 16, 16), (96, 64, 80), (8, 8, 8), (96, 64, 80)):
        b = scenes.noise_volume_scene(res=res, density=20.0, light="rect")
        means.append(float(frame(ctx, b, cam, 32, 24, 16).mean()))
    assert all(np.isfinite(means)) and abs(means[1] - means[3]) / means[1] < 1e-3   # same scene, same seed, same image
