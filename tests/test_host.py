"""CPU tests of the host side of the boundary: glm-order transform / camera arithmetic against the oracle, the scene
builders (brick-sparse grid, binned-SAH BVH) against brute force in numpy, the Philox restatement against the published
known-answer vectors, and the sample-index partition over a world_size-2 gloo group."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import scenes
from narvalengine_b200 import abi
from narvalengine_b200.multigpu import PartitionedFrame, sample_range
from refclient import RefOracle


@pytest.fixture(scope="module")
def oracle():
    return RefOracle()


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()


# ---------------------------------------------------------------------------------------------------------------
def test_philox_restatement_matches_random123_kat():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    from philox_ref import philox4x32_10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_make_transform_is_bit_exact_with_glm(lib, oracle):
    """ne_b200_make_transform restates getTransform + glm::inverse (Math.h:848-859, InstancedModel.cpp:11-22)."""
    rng = np.random.default_rng(2)
    cases = [((1, 2, 3), (30, 45, 60), (2, 3, 4)), ((0, 3.9, 0), (-89, 0, 0), (1, 1, 1)), ((0, 0, 0), (90, 0, 0), (4, 4, 1))]
    cases += [(rng.uniform(-5, 5, 3), rng.uniform(-180, 180, 3), rng.uniform(.1, 6, 3)) for _ in range(200)]
    for p, r, s in cases:
        M, Mi = (C.c_float * 16)(), (C.c_float * 16)()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])  # noqa: E731
        assert lib.ne_b200_make_transform(f3(p), f3(r), f3(s), M, Mi) == 0
        rM, rMi = oracle.get_transform(p, r, s)
        assert np.array_equal(np.array(M, np.float32), rM), (p, r, s)
        assert np.array_equal(np.array(Mi, np.float32), rMi), (p, r, s)


def test_camera_make_is_bit_exact_with_the_reference(lib, oracle):
    rng = np.random.default_rng(3)
    cams = [scenes.CORNELL_CAMERA, scenes.MESH_CAMERA, scenes.MIXED_CAMERA]
    cams += [scenes.CameraParams(rng.uniform(-6, 6, 3), rng.uniform(-1, 1, 3), float(rng.uniform(20, 90))) for _ in range(50)]
    for cp in cams:
        for aspect in (1.0, 16 / 9):
            a, b = cp.make(aspect, lib), oracle.camera_make(cp, aspect)
            for f, _ in abi.Camera._fields_:
                va, vb = getattr(a, f), getattr(b, f)
                assert (va == vb) if f == "lens_radius" else (list(va) == list(vb)), f


def test_null_arguments_are_rejected(lib):
    assert lib.ne_b200_make_transform(None, None, None, None, None) == abi.ERR_INVALID
    assert lib.ne_b200_camera_make(None, None, None, 45.0, 1.0, 0.0, 1.0, None) == abi.ERR_INVALID
    assert lib.ne_b200_host_build_bricks(None, None, None, None, None, None) == abi.ERR_INVALID
    assert lib.ne_b200_host_build_bvh(None, 0, None, 3, None, None, None) == abi.ERR_INVALID


# ---------------------------------------------------------------------------------------------------------------
# Brick-sparse grid
# ---------------------------------------------------------------------------------------------------------------
def build_bricks(lib, vol):
    dims = (C.c_int32 * 4)()
    mx = C.c_float()
    assert lib.ne_b200_host_build_bricks(C.byref(vol), dims, None, None, None, C.byref(mx)) == 0
    bx, by, bz, slots = list(dims)
    table = np.zeros((bz, by, bx), np.int32)
    binv = np.zeros((bz, by, bx), np.float32)
    pool = np.zeros((max(slots, 1), 9, 9, 9), np.float32)
    assert lib.ne_b200_host_build_bricks(C.byref(vol), dims, table.ctypes.data_as(abi.pi32), binv.ctypes.data_as(abi.pf32),
                                         pool.ctypes.data_as(abi.pf32), C.byref(mx)) == 0
    return table, binv, pool[:slots], mx.value


def check_bricks(grid, table, binv, pool, mx):
    D, H, W = grid.shape
    assert mx == grid.max()
    pad = np.zeros((table.shape[0] * 8 + 9, table.shape[1] * 8 + 9, table.shape[2] * 8 + 9), np.float32)  # zeros beyond the grid
    pad[:D, :H, :W] = grid
    used = set()
    for bz, by, bx in np.ndindex(*table.shape):
        region9 = pad[bz * 8:bz * 8 + 9, by * 8:by * 8 + 9, bx * 8:bx * 8 + 9]
        s = table[bz, by, bx]
        if s < 0:
            assert not region9.any(), (bx, by, bz)  # no storage only where every reachable voxel is zero
        else:
            assert s not in used
            used.add(int(s))
            assert np.array_equal(pool[s], region9), (bx, by, bz)
        lo = [max(0, 8 * b - 1) for b in (bz, by, bx)]
        support = pad[lo[0]:bz * 8 + 9, lo[1]:by * 8 + 9, lo[2]:bx * 8 + 9]  # [8b-1, 8b+8]^3: every trilinear stencil of the brick
        m = support.max()
        assert binv[bz, by, bx] == (np.float32(1.0) / m if m > 0 else 0.0), (bx, by, bz)
    assert used == set(range(len(pool)))


@pytest.mark.parametrize("shape", [(4, 4, 4), (16, 16, 16), (13, 21, 9), (40, 24, 32)])
def test_brick_builder_dense(lib, shape):
    rng = np.random.default_rng(sum(shape))
    g = rng.uniform(0, 1, shape).astype(np.float32)
    g[rng.uniform(0, 1, shape) < 0.6] = 0
    g[: shape[0] // 2] = 0  # whole empty bricks
    b = scenes.SceneBuilder(transform_fn=lambda p, r, s: ([0] * 16, [0] * 16))
    b.add_volume_dense(g)
    check_bricks(g, *build_bricks(lib, b.volumes[0]))


def test_brick_builder_leaves_equal_dense(lib):
    g = scenes.cloud_density((40, 24, 32), seed=7)  # [z=32? no: res is (W,H,D)] -> array [D,H,W]
    o, v = scenes.dense_to_leaves(g)
    D, H, W = g.shape
    b = scenes.SceneBuilder(transform_fn=lambda p, r, s: ([0] * 16, [0] * 16))
    b.add_volume_dense(g)
    b.add_volume_leaves((W, H, D), o, v)
    td, bd, pd, md = build_bricks(lib, b.volumes[0])
    tl, bl, pl, ml = build_bricks(lib, b.volumes[1])
    assert md == ml and np.array_equal(td, tl) and np.array_equal(bd, bl) and np.array_equal(pd, pl)
    check_bricks(g, tl, bl, pl, ml)


def test_brick_builder_edge_cases(lib):
    b = scenes.SceneBuilder(transform_fn=lambda p, r, s: ([0] * 16, [0] * 16))
    b.add_volume_dense(np.zeros((8, 8, 8), np.float32))                       # all empty
    b.add_volume_leaves((16, 16, 16), np.zeros((0, 3), np.int32), np.zeros((0, 8, 8, 8), np.float32))  # no leaves at all
    one = np.zeros((1, 8, 8, 8), np.float32)
    one[0, 7, 7, 7] = 2.5                                                      # one voxel in the corner of leaf (8,8,8)
    b.add_volume_leaves((16, 16, 16), [[8, 8, 8]], one)
    for i in (0, 1):
        t, binv, pool, mx = build_bricks(lib, b.volumes[i])
        assert (t < 0).all() and (binv == 0).all() and len(pool) == 0 and mx == 0
    t, binv, pool, mx = build_bricks(lib, b.volumes[2])
    g = np.zeros((16, 16, 16), np.float32)
    g[15, 15, 15] = 2.5
    check_bricks(g, t, binv, pool, mx)
    assert (t >= 0).sum() == 1 and mx == 2.5


# ---------------------------------------------------------------------------------------------------------------
# BVH
# ---------------------------------------------------------------------------------------------------------------
NODE = np.dtype([("lo0", "<f4", 3), ("hi0", "<f4", 3), ("lo1", "<f4", 3), ("hi1", "<f4", 3), ("child", "<i4", 2), ("cnt", "<i4", 2)])


def build_bvh(lib, pos, idx):
    pos, idx = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(idx, np.uint32)
    counts = (C.c_int32 * 2)()
    args = (pos.ctypes.data_as(abi.pf32), len(pos), idx.ctypes.data_as(abi.pu32), len(idx))
    assert lib.ne_b200_host_build_bvh(*args, counts, None, None) == 0
    nodes = np.zeros(counts[0], NODE)
    tris = np.zeros((counts[1], 12), np.float32)
    assert lib.ne_b200_host_build_bvh(*args, counts, nodes.ctypes.data_as(C.c_void_p), tris.ctypes.data_as(abi.pf32)) == 0
    return nodes, tris


def walk(nodes, tris, ref, lo, hi, seen):
    """Every triangle under `ref` lies inside [lo, hi]; returns nothing, fills `seen` with original triangle ids."""
    if ref < 0:
        enc = ~ref
        first, cnt = enc >> 3, enc & 7
        assert 0 < cnt <= 4
        for s in range(first, first + cnt):
            v = tris[s].reshape(3, 4)[:, :3]
            assert (v >= lo - 1e-6).all() and (v <= hi + 1e-6).all()
            seen.append(int(tris[s, 3:4].view(np.int32)[0]))
        return
    n = nodes[ref]
    for k in (0, 1):
        clo, chi = n[f"lo{k}"], n[f"hi{k}"]
        assert (clo >= lo - 1e-6).all() and (chi <= hi + 1e-6).all()  # child boxes nest
        walk(nodes, tris, int(n["child"][k]), clo, chi, seen)


@pytest.mark.parametrize("n", [1, 2, 8, 24])
def test_bvh_builder_structure(lib, n):
    pos, idx, _ = scenes.displaced_grid(n)
    nodes, tris = build_bvh(lib, pos, idx)
    assert len(tris) == len(idx)
    seen = []
    if len(idx) <= 4:  # the whole mesh is one leaf under a root whose second child is empty
        walk(nodes, tris, int(nodes[0]["child"][0]), nodes[0]["lo0"], nodes[0]["hi0"], seen)
    else:
        walk(nodes, tris, 0, pos.min(0), pos.max(0), seen)
    assert sorted(seen) == list(range(len(idx)))  # every triangle exactly once
    for s in range(len(tris)):  # slots hold the vertices of the triangle they name
        t = int(tris[s, 3:4].view(np.int32)[0])
        assert np.array_equal(tris[s].reshape(3, 4)[:, :3], pos[idx[t]])


def test_bvh_builder_degenerate_inputs(lib):
    nodes, tris = build_bvh(lib, np.zeros((3, 3), np.float32), np.zeros((0, 3), np.uint32))  # no triangles
    assert len(tris) == 0 and len(nodes) == 1
    pos = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (1, 1))
    idx = np.tile(np.array([[0, 1, 2]], np.uint32), (37, 1))  # 37 coincident triangles: centroid extent 0 -> median split
    nodes, tris = build_bvh(lib, pos, idx)
    seen = []
    walk(nodes, tris, 0, pos.min(0), pos.max(0), seen)
    assert sorted(seen) == list(range(37))
    counts = (C.c_int32 * 2)()
    bad = np.array([[0, 1, 7]], np.uint32)
    assert lib.ne_b200_host_build_bvh(pos.ctypes.data_as(abi.pf32), 3, bad.ctypes.data_as(abi.pu32), 1, counts, None, None) == abi.ERR_INVALID


# ---------------------------------------------------------------------------------------------------------------
# Scene builder / fold order
# ---------------------------------------------------------------------------------------------------------------
def test_scene_builder_fold_order_matches_the_reference(oracle):
    """SceneReader puts emitter primitives in Scene::lights and the rest in Scene::instancedModels (JSON order each);
    the oracle builds the reference's own Scene from the same descriptor."""
    for mk, want in ((scenes.s1_cornell, (6, 1)), (scenes.s2_volume, (1, 1)), (scenes.cornell_c1, (5, 2)), (scenes.mixed_scene, (4, 2))):
        rs = oracle.scene(mk())
        assert rs.counts() == want
        rs.close()


# ---------------------------------------------------------------------------------------------------------------
# Sample-index partition (multi-GPU path) on CPU: world_size 2, gloo
# ---------------------------------------------------------------------------------------------------------------
def test_sample_range_tiles_the_sample_axis():
    for spp in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            r = [sample_range(k, world, spp) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sample_range(2, 2, 8)


class FakeContext:
    """Stands in for engine.Context on CPU: 'renders' a deterministic value per (pixel, sample) from the Philox
    restatement, so partitioned and whole-frame sums can be compared exactly (up to fp32 order)."""

    def __init__(self, torch, n_pix):
        self.torch, self.n_pix = torch, n_pix
        self.accum = torch.zeros(n_pix * 3, dtype=torch.float32)
        self.samples = 0

    def clear(self):
        self.accum.zero_()
        self.samples = 0

    def render(self, W, H, begin, end, bounces, seed=1, flags=0):
        from philox_ref import philox_uniforms
        for px in range(self.n_pix):
            for s in range(begin, end):
                u = philox_uniforms(seed, px, s, 3)
                self.accum[3 * px:3 * px + 3] += self.torch.from_numpy(u)
        self.samples += end - begin

    def set_samples_accumulated(self, n):
        self.samples = n


def _worker(rank, world, port, spp, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = FakeContext(torch, 6)
        frame = PartitionedFrame(ctx, ctx.accum, rank, world, dist)
        b, e = frame.render(3, 2, spp, 6, seed=9)
        dist.barrier()
        if rank == 0:
            np.save(out, np.concatenate([ctx.accum.numpy(), [ctx.samples, b, e]]))
    finally:
        dist.destroy_process_group()


def test_partitioned_frame_over_gloo_equals_single_process(tmp_path):
    import torch
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    spp, out = 5, str(tmp_path / "accum.npy")
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    got = np.load(out)
    single = FakeContext(torch, 6)
    PartitionedFrame(single, single.accum).render(3, 2, spp, 6, seed=9)
    assert got[-3] == spp and (got[-2], got[-1]) == (0, 3)  # rank 0 rendered samples [0,3), rank 1 [3,5)
    np.testing.assert_allclose(got[:-3], single.accum.numpy(), rtol=1e-6)
    assert single.samples == spp
