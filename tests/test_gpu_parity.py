"""GPU parity tests proper: every device function of the hot path, called through the C ABI (include/ne_b200.h), against
the oracle (the reference's own translation units, oracle/_ref/libnarval_ref.so) on the same seeded inputs.

Tolerance: 1e-5 relative (BASELINE.json north_star) for per-function values; tape tests feed the CUDA side the exact
uniform sequence narvalengine::random() produced on the oracle side."""
import numpy as np
import pytest

import scenes
from narvalengine_b200 import abi
from narvalengine_b200.engine import Context
from refclient import RefOracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def oracle():
    return RefOracle()


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def close(a, b, rtol=RTOL, atol=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    with np.errstate(invalid="ignore"):
        return (np.abs(a - b) <= atol + rtol * np.abs(b)) | (np.isnan(a) & np.isnan(b)) | (a == b)


def assert_close(a, b, rtol=RTOL, atol=1e-6, frac=1.0, what=""):
    ok = close(a, b, rtol, atol)
    if ok.mean() < frac:
        bad = np.argwhere(~ok)[:5]
        raise AssertionError(f"{what}: only {ok.mean():.6f} within tol; first bad {bad.tolist()} got {np.asarray(a)[tuple(bad[0])]} "
                             f"want {np.asarray(b)[tuple(bad[0])]}")


def hits_to_arrays(hits, n):
    d = dict(hit=np.zeros(n, np.int32), inst=np.zeros(n, np.int32), prim=np.zeros(n, np.int32), t=np.zeros((n, 2), np.float32),
             p=np.zeros((n, 3), np.float32), nrm=np.zeros((n, 3), np.float32), uv=np.zeros((n, 2), np.float32), light=np.zeros(n, np.int32))
    for i in range(n):
        h = hits[i]
        d["hit"][i], d["inst"][i], d["prim"][i], d["light"][i] = h.hit, h.instance, h.primitive, h.is_light
        d["t"][i] = (h.t_near, h.t_far)
        d["p"][i], d["nrm"][i], d["uv"][i] = tuple(h.hit_point), tuple(h.normal), tuple(h.uv)
    return d


def compare_hits(ctx, ref_scene, o, d, tmin=1e-11, min_frac=1.0, what=""):
    n = len(o)
    a = hits_to_arrays(ctx.intersect(o, d, tmin), n)
    b = hits_to_arrays(ref_scene.intersect(o, d, tmin), n)
    same = (a["hit"] == b["hit"]) & (a["inst"] == b["inst"])
    assert same.mean() >= min_frac, f"{what}: hit/instance agreement {same.mean()}"
    m = same & (b["hit"] == 1)
    assert m.sum() > 0, what
    for k in ("t", "p", "nrm", "uv"):
        assert_close(a[k][m], b[k][m], what=f"{what}.{k}", atol=2e-6)
    assert (a["light"][m] == b["light"][m]).all()
    return a, b, m


def random_rays(n, seed, center=(0, 2, 0), spread=1.5):
    rng = np.random.default_rng(seed)
    o = (np.asarray(center) + rng.uniform(-spread, spread, (n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return o, d


# ---------------------------------------------------------------------------------------------------------------
def test_philox_matches_published_algorithm(ctx):
    from philox_ref import philox_uniforms
    for seed, px, smp in ((1, 0, 0), (0xDEADBEEF12345678, 123456, 63), (7, 2073599, 255)):
        got = ctx.philox(seed, px, smp, 64)
        want = philox_uniforms(seed, px, smp, 64)
        assert np.array_equal(got, want)


def test_camera_rays(ctx, oracle):
    for aspect in (1.0, 1920 / 1080):
        cam = scenes.CORNELL_CAMERA.make(aspect, ctx.lib)
        ctx.set_camera(cam)
        rng = np.random.default_rng(3)
        xy = rng.uniform(0, 1, (256, 2)).astype(np.float32)
        xy[:4] = [(0, 0), (1, 0), (0, 1), (.25, .75)]
        ro, rd = oracle.camera_rays(scenes.CORNELL_CAMERA, aspect, 123, xy)
        tape = oracle.tape(123, 2 * len(xy))
        o, d = ctx.camera_rays(xy, tape)
        assert_close(o, ro, atol=1e-7, what="camera.o")
        assert_close(d, rd, atol=1e-6, what="camera.d")


def test_intersect_cornell(ctx, oracle):
    b = scenes.s1_cornell()
    ctx.upload(b)
    rs = oracle.scene(b)
    o, d = random_rays(20000, 1)
    compare_hits(ctx, rs, o, d, what="S1 random")
    # shadow-ray convention: unnormalised directions, tMin 1e-3
    compare_hits(ctx, rs, o, d * 3.7, tmin=1e-3, what="S1 shadow")
    # rays leaving the sphere surface (Q14: the 'inside' branch is decided by rounding)
    rng = np.random.default_rng(5)
    n = rng.normal(size=(5000, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    po = (np.array([.6, .6, .5]) + .6 * n).astype(np.float32)
    dd = n + rng.normal(size=n.shape) * .7
    dd = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    compare_hits(ctx, rs, po, dd, what="S1 sphere self-hit", min_frac=0.999)


def test_golden_hits_appendix_e(ctx, oracle):
    """SURVEY Appendix E9/E10/E14."""
    ctx.upload(scenes.s1_cornell())
    d9 = np.array([.1, -.25, 1.]) / np.linalg.norm([.1, -.25, 1.])
    d10 = np.array([-.3, .1, 1.]) / np.linalg.norm([-.3, .1, 1.])
    h = ctx.intersect([(0, 2, -5), (0, 2, -5)], [d9, d10])
    assert h[0].instance == 5 and abs(h[0].t_near - 5.10925674) < 1e-4
    assert_close(tuple(h[0].normal), (-0.177741334, 0.277687728, -0.944085598), atol=1e-5)
    assert h[1].instance == 3 and abs(h[1].t_near - 6.99205875) < 1e-4
    assert_close(tuple(h[1].uv), (0.916666627, 0.666666627), atol=1e-5)
    ctx.upload(scenes.s2_volume())
    h = ctx.intersect([(-.4, 1.3, -5)], [(0, 0, 1)])
    assert h[0].hit and abs(h[0].t_near - 4) < 1e-5 and abs(h[0].t_far - 6) < 1e-5
    assert_close(tuple(h[0].normal), (0, 0, -1))


def test_intersect_volume_and_mixed(ctx, oracle):
    for sort in (False, True):
        b = scenes.mixed_scene(sort_and_group=sort)
        ctx.upload(b)
        rs = oracle.scene(b)
        o, d = random_rays(20000, 11 + sort, center=(0, 1.5, 0), spread=2.5)
        compare_hits(ctx, rs, o, d, what=f"mixed sort={sort}")
        compare_hits(ctx, rs, o, d * 2.5, tmin=1e-3, what=f"mixed shadow sort={sort}")


def test_intersect_mesh(ctx, oracle):
    b = scenes.mesh_scene(n=48)
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(2)
    n = 20000
    o = np.stack([rng.uniform(-2, 2, n), rng.uniform(0.5, 3, n), rng.uniform(-2, 2, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d[:, 1] = -np.abs(d[:, 1])
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a, r, m = compare_hits(ctx, rs, o, d, what="mesh", min_frac=0.9995)
    assert (a["prim"][m] == r["prim"][m]).mean() > 0.9995
    # secondary rays starting ON the mesh (Q6 self-hit semantics)
    po = r["p"][m][:5000]
    dd = rng.normal(size=po.shape)
    dd[:, 1] = np.abs(dd[:, 1])
    dd = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    compare_hits(ctx, rs, po, dd, what="mesh secondary", min_frac=0.999)
    compare_hits(ctx, rs, po, dd, tmin=1e-3, what="mesh secondary shadow", min_frac=0.999)


def test_reference_unit_test_vectors(ctx, oracle):
    """unitTests/tests.cpp Triangle.intersection :300-347 and Model.modelMadeOfTriangles :409-481 restated on a
    one-triangle / cube mesh (t only; the mesh goes through our BVH)."""
    b = scenes.SceneBuilder()
    b.add_microfacet("m", (.5, .5, .5), .5, 0)
    b.add_emitter("l", (1, 1, 1))
    b.add_mesh("m", [(0, 0, 0), (0, 1, 0), (1, 0, 0)], [(0, 1, 2)])
    b.add_rectangle("l", (0, 50, 0))
    ctx.upload(b)
    o = [(0, 0, -1), (0, 0, -1), (.25, .25, -1), (.25, .25, 1)]
    d = [(0, 0, 1), (1, 0, 0), (0, 0, 1), (0, 0, 1)]
    h = ctx.intersect(o, d, tmin=-1.0)
    assert h[0].hit and abs(h[0].t_near - 1) < 1e-6       # hit at a vertex
    assert not h[1].hit                                     # parallel
    assert h[2].hit and abs(h[2].t_near - 1) < 1e-6       # centre
    assert not h[3].hit                                     # behind
    rs = oracle.scene(b)
    r = rs.intersect(o, d, tmin=-1.0)
    assert [x.hit for x in r] == [x.hit for x in h]


def test_bsdf(ctx, oracle):
    b = scenes.s1_cornell()
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(4)
    n = 4000
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    inc = rng.normal(size=(n, 3)); inc *= rng.uniform(.5, 3, (n, 1))
    sc = rng.normal(size=(n, 3)); sc /= np.linalg.norm(sc, axis=1, keepdims=True)
    flip = np.sum(inc * nrm, 1) > 0
    inc[flip] *= -1  # mostly valid configurations, some invalid ones are kept below
    inc[:200] *= -1
    seeds = np.arange(n, dtype=np.uint32) + 100
    tape = np.stack([oracle.tape(int(s), 2) for s in seeds])
    for inst in (0, 3, 5):  # white wall (rough .95), red wall, sphere (rough .5)
        re, rp, rsmp = rs.bsdf(inst, inc, sc, nrm, seeds=seeds)
        e, p, smp = ctx.bsdf(inst, inc, sc, nrm, tape=tape)
        assert_close(e, re, what=f"eval[{inst}]", atol=1e-7)
        assert_close(p, rp, what=f"pdf[{inst}]", atol=1e-7)
        assert_close(smp, rsmp, what=f"sample[{inst}]", rtol=2e-5, atol=2e-6)
    # SURVEY E4
    bb = scenes.SceneBuilder()
    bb.add_microfacet("m", (.8, .6, .4), .5, .2)
    bb.add_emitter("l", (1, 1, 1))
    bb.add_rectangle("m", (0, 0, 0))
    bb.add_rectangle("l", (0, 5, 0))
    ctx.upload(bb)
    i = np.array([1, -1, .5]) / np.linalg.norm([1, -1, .5])
    s = np.array([-.3, .8, .2]) / np.linalg.norm([-.3, .8, .2])
    e, p, _ = ctx.bsdf(0, [i], [s], [(0, 1, 0)])
    assert_close(e[0], (0.180891529, 0.136317134, 0.0917427093), atol=1e-7)
    assert_close(p[0], 0.0394177698, atol=1e-7)


def test_bsdf_image_texture(ctx, oracle):
    b = scenes.textured_scene()
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(8)
    n = 2000
    nrm = np.tile([0, 1, 0], (n, 1)).astype(np.float32)
    inc = rng.normal(size=(n, 3)); inc[:, 1] = -np.abs(inc[:, 1])
    sc = rng.normal(size=(n, 3)); sc[:, 1] = np.abs(sc[:, 1])
    uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
    re, rp, _ = rs.bsdf(0, inc, sc, nrm, uvs=uv)
    e, p, _ = ctx.bsdf(0, inc, sc, nrm, uvs=uv)
    assert_close(e, re, what="textured eval", atol=1e-7)
    assert_close(p, rp, what="textured pdf", atol=1e-7)


def test_phase_function(ctx, oracle):
    for phase, g in (("hg", 0.0), ("hg", 0.7), ("hg", -0.3), ("isotropic", 0.0)):
        b = scenes.s2_volume(phase=phase, g=g)
        ctx.upload(b)
        rs = oracle.scene(b)
        rng = np.random.default_rng(6)
        n = 2000
        nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        inc = rng.normal(size=(n, 3))
        sc = rng.normal(size=(n, 3)) * 2
        seeds = np.arange(n, dtype=np.uint32) + 7
        tape = np.stack([oracle.tape(int(s), 2) for s in seeds])
        re, rp, rsmp = rs.bsdf(0, inc, sc, nrm, seeds=seeds)
        e, p, smp = ctx.bsdf(0, inc, sc, nrm, tape=tape)
        assert_close(e, re, what=f"phase eval {phase} {g}", atol=1e-8)
        assert_close(p, rp, what=f"phase pdf {phase} {g}", atol=1e-8)
        assert_close(smp, rsmp, what=f"phase sample {phase} {g}", rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("leaves", [False, True])
def test_density_and_tracking(ctx, oracle, leaves):
    b = scenes.noise_volume_scene(res=(40, 24, 33), leaves=leaves, density=30.0)
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(9)
    pts = rng.uniform(-.5, .5, (20000, 3)).astype(np.float32)
    pts[:64] = np.where(rng.uniform(size=(64, 3)) < .5, -.5, .5)  # faces, edges, corners
    pts[64:128] = rng.uniform(-.6, .6, (64, 3))                   # slightly outside: clamped
    rd, rinv = rs.density(0, pts)
    dd, inv = ctx.density(0, pts)
    assert inv == pytest.approx(rinv, rel=1e-7)
    assert np.array_equal(dd, rd), f"max diff {np.abs(dd - rd).max()}"

    # rays through the volume: take the hit record from the scene fold, then Tr / sample with the oracle's tape
    n = 3000
    o = np.stack([rng.uniform(-3, 3, n), rng.uniform(-1, 3, n), np.full(n, -6.0)], 1).astype(np.float32)
    tgt = np.stack([rng.uniform(-1.2, 1.2, n), 1 + rng.uniform(-1.2, 1.2, n), rng.uniform(-1.2, 1.2, n)], 1)
    d = (tgt - o)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(.5, 2.5, (n, 1))).astype(np.float32)
    hits = hits_to_arrays(rs.intersect(o, d), n)
    m = hits["hit"] == 1
    o, d, tn, tf = o[m], d[m], hits["t"][m, 0], hits["t"][m, 1]
    n = len(o)
    assert n > 1000
    seeds = np.arange(n, dtype=np.uint32) + 1000
    stride = 4096
    tape = np.stack([oracle.tape(int(s), stride) for s in seeds])
    rtr, rused = rs.grid_tr(0, o, d, tn, tf, seeds)
    tr, used = ctx.grid_tr(0, o, d, tn, tf, tape)
    same = used == rused
    assert same.mean() > 0.995, f"Tr draw-count agreement {same.mean()}"
    assert_close(tr[same], rtr[same], what="Tr", rtol=1e-4, atol=1e-6, frac=0.999)

    # sample(): called the way Li calls it (origin moved to the entry point, tNear = 0)
    o2 = (o + tn[:, None] * d).astype(np.float32)
    z = np.zeros(n, np.float32)
    rT, rso, rsd, rused = rs.grid_sample(0, o2, d, z, tf - tn, seeds)
    T, so, sd, used = ctx.grid_sample(0, o2, d, z, tf - tn, tape)
    same = used == rused
    assert same.mean() > 0.995, f"sample draw-count agreement {same.mean()}"
    assert_close(T[same], rT[same], what="sample.T")
    coll = same & (rused > 0) & ~np.all(rT == 1, axis=1)
    assert coll.sum() > 100
    assert_close(so[coll], rso[coll], what="sample.o", rtol=1e-4, atol=1e-4, frac=0.999)
    assert_close(sd[coll], rsd[coll], what="sample.d", rtol=1e-4, atol=1e-5, frac=0.999)


@pytest.mark.parametrize("res", [(40, 24, 33), (64, 64, 64), (9, 8, 7)])
def test_gpu_brick_builder_equals_host_builder(ctx, res):
    """Dense grids are bricked by GPU kernels during ne_b200_scene_upload (csrc/ne_bricks.cu); the result must be the
    host builder's (ne_b200_host_build_bricks, the inspectable definition checked voxel by voxel in test_host.py):
    same slot order, same records, same reciprocal majorants, same global maximum - bit for bit."""
    import ctypes as C
    b = scenes.noise_volume_scene(res=res, density=30.0)
    ctx.upload(b)
    lib = ctx.lib
    vol = b.desc().volumes[0]
    dims = (C.c_int32 * 4)()
    mx = C.c_float()
    assert lib.ne_b200_host_build_bricks(C.byref(vol), dims, None, None, None, C.byref(mx)) == 0
    bx, by, bz, slots = list(dims)
    ht, hi, hp = np.zeros((bz, by, bx), np.int32), np.zeros((bz, by, bx), np.float32), np.zeros((max(slots, 1), 729), np.float32)
    assert lib.ne_b200_host_build_bricks(C.byref(vol), dims, ht.ctypes.data_as(abi.pi32), hi.ctypes.data_as(abi.pf32),
                                         hp.ctypes.data_as(abi.pf32), C.byref(mx)) == 0
    ddims = (C.c_int32 * 4)()
    dmx = C.c_float()
    assert lib.ne_b200_test_read_bricks(ctx.h, 0, ddims, None, None, None, C.byref(dmx)) == 0
    assert list(ddims) == [bx, by, bz, slots] and dmx.value == mx.value
    dt, di, dp = np.zeros_like(ht), np.zeros_like(hi), np.zeros_like(hp)
    assert lib.ne_b200_test_read_bricks(ctx.h, 0, ddims, dt.ctypes.data_as(abi.pi32), di.ctypes.data_as(abi.pf32),
                                        dp.ctypes.data_as(abi.pf32), None) == 0
    assert np.array_equal(dt, ht)
    assert np.array_equal(di.view(np.uint32), np.where(ht >= 0, hi, 0).astype(np.float32).view(np.uint32))
    assert np.array_equal(dp[:slots].view(np.uint32), hp[:slots].view(np.uint32))


def test_golden_tracking_appendix_e(ctx, oracle):
    """SURVEY E15 (Tr, seed 5: killed by RR after 5 draws) and E17 (sample, seed 6)."""
    ctx.upload(scenes.s2_volume())
    tr, used = ctx.grid_tr(0, [(-.4, 1.3, -5)], [(0, 0, 1)], [4], [6], oracle.tape(5, 64))
    assert tr[0] == 0 and used[0] == 5
    T, so, sd, used = ctx.grid_sample(0, [(-.4, 1.3, -1)], [(0, 0, 1)], [0], [2], oracle.tape(6, 64))
    assert used[0] == 6
    assert_close(T[0], (0.990990996,) * 3)
    assert_close(so[0], (-0.400000006, 1.29999995, -0.208090901), atol=1e-6)
    assert_close(sd[0], (0.884224415, 1.74298155, -0.424455553), atol=1e-5)


def _tape_li(ctx, oracle, rs, o, d, bounces, seeds, stride=8192):
    tape = np.stack([oracle.tape(int(s), stride) for s in seeds])
    rL, rused = rs.li(o, d, bounces, seeds)
    L, used = ctx.li_tape(o, d, bounces, tape)
    return L, used, rL, rused


def test_sample_one_light_tape(ctx, oracle):
    b = scenes.s1_cornell(with_sphere=False)
    ctx.upload(b)
    rs = oracle.scene(b)
    o, d = random_rays(4000, 21, center=(0, 2, -1), spread=.8)
    hits = rs.intersect(o, d)
    n = len(o)
    keep = [i for i in range(n) if hits[i].hit and not hits[i].is_light]
    hh = (abi.Hit * len(keep))(*[hits[i] for i in keep])
    dirs = d[keep]
    seeds = np.arange(len(keep), dtype=np.uint32) + 300
    tape = np.stack([oracle.tape(int(s), 256) for s in seeds])
    rL, rused = rs.sample_one_light(dirs, hh, seeds)
    L, used = ctx.sample_one_light(dirs, hh, tape)
    same = used == rused
    assert same.mean() > 0.998, f"draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="uniformSampleOneLight", rtol=5e-5, atol=1e-6, frac=0.998)
    assert (rL.sum(axis=1) > 0).mean() > 0.2


def test_li_tape_surfaces(ctx, oracle):
    b = scenes.s1_cornell(with_sphere=False)
    ctx.upload(b)
    rs = oracle.scene(b)
    cam_o, cam_d = oracle.camera_rays(scenes.CORNELL_CAMERA, 1.0, 99, np.random.default_rng(1).uniform(0, 1, (3000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 5000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds)
    same = used == rused
    assert same.mean() > 0.99, f"Li draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li surfaces", rtol=1e-4, atol=1e-5, frac=0.995)
    assert rL.mean() > 1


def test_li_tape_golden_appendix_e(ctx, oracle):
    """SURVEY E12/E13 (S1 with the sphere) and E18 (S2)."""
    ctx.upload(scenes.s1_cornell())
    d9 = np.array([.1, -.25, 1.]) / np.linalg.norm([.1, -.25, 1.])
    d10 = np.array([-.3, .1, 1.]) / np.linalg.norm([-.3, .1, 1.])
    L, used = ctx.li_tape([(0, 2, -5)], [d9], 6, oracle.tape(11, 4096))
    assert used[0] == 16
    assert_close(L[0], (0.27675885, 0.27675885, 0.354118615), rtol=1e-4)
    L, used = ctx.li_tape([(0, 2, -5)], [d10], 6, oracle.tape(12, 4096))
    assert used[0] == 19
    assert_close(L[0], (3.22971058, 0.255419731, 0.369815588), rtol=1e-4)
    ctx.upload(scenes.s2_volume())
    L, used = ctx.li_tape([(-.4, 1.3, -5)], [(0, 0, 1)], 6, oracle.tape(21, 4096))
    assert used[0] == 32
    assert_close(L[0], (0.00700760074,) * 3, rtol=1e-4)


def test_li_tape_volume(ctx, oracle):
    b = scenes.noise_volume_scene(res=(24, 24, 24), density=12.0, light="rect")
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(31)
    n = 1500
    o = np.tile([0, 1, -6], (n, 1)).astype(np.float32)
    tgt = np.stack([rng.uniform(-1, 1, n), 1 + rng.uniform(-1, 1, n), np.zeros(n)], 1)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    seeds = np.arange(n, dtype=np.uint32) + 9000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, o, d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"Li(volume) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li volume", rtol=2e-4, atol=1e-6, frac=0.99)
    assert (rL.sum(axis=1) > 0).mean() > 0.05


def test_li_tape_mixed(ctx, oracle):
    b = scenes.mixed_scene(sort_and_group=True)
    ctx.upload(b)
    rs = oracle.scene(b)
    cp = scenes.MIXED_CAMERA
    cam_o, cam_d = oracle.camera_rays(cp, 1.0, 5, np.random.default_rng(2).uniform(0, 1, (2000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 70000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"Li(mixed) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li mixed", rtol=2e-4, atol=1e-5, frac=0.99)


def test_li_tape_directional_light(ctx, oracle):
    """DirectionalLight (lights/DirectionalLight.cpp, Q23): Le on camera rays, a dead light half that draws nothing, Li()
    through the BSDF half only - whole Li paths draw for draw against the reference."""
    b = scenes.directional_scene()
    ctx.upload(b)
    rs = oracle.scene(b)
    cam_o, cam_d = oracle.camera_rays(scenes.DIRECTIONAL_CAMERA, 1.0, 7, np.random.default_rng(4).uniform(0, 1, (2000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 90000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"Li(directional) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li directional", rtol=2e-4, atol=1e-5, frac=0.99)
    assert (rL.min(axis=1) >= 0.4 - 1e-6).mean() > 0.9  # nearly every camera ray carries the sun's Le


def test_li_tape_infinite_area_light(ctx, oracle):
    """InfiniteAreaLight + Distribution1D/2D (lights/InfiniteAreaLight.h, utils/Sampling.h incl. Q27): Le on camera rays,
    importance-sampled light half, never-hit carrier sphere - whole Li paths draw for draw against the reference."""
    b = scenes.environment_scene()
    ctx.upload(b)
    rs = oracle.scene(b)
    cam_o, cam_d = oracle.camera_rays(scenes.ENVIRONMENT_CAMERA, 1.0, 7, np.random.default_rng(4).uniform(0, 1, (2000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 90000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"Li(environment) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li environment", rtol=2e-4, atol=1e-5, frac=0.99)
    assert (rused > 0).mean() > 0.3 and rL.mean() > 0.5


@pytest.mark.parametrize("grey", [True, False])
def test_li_tape_homogeneous_media(ctx, oracle, grey):
    """HomogeneousMedia (materials/HomogeneousMedia.cpp:15-51): closed-form sample with its t / distance quirk, the
    rounding-dependent escape detection, vec3 transmittance through intersectTr - whole Li paths draw for draw."""
    b = scenes.homogeneous_scene(grey=grey)
    ctx.upload(b)
    rs = oracle.scene(b)
    cam_o, cam_d = oracle.camera_rays(scenes.HOMOGENEOUS_CAMERA, 1.0, 7, np.random.default_rng(6).uniform(0, 1, (2000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 110000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"Li(homogeneous) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li homogeneous", rtol=2e-4, atol=1e-5, frac=0.99)
    assert (rused > 0).mean() > 0.2


# ---------------------------------------------------------------------------------------------------------------
# The light of BASELINE configs[1]: a DiffuseLight on a Point primitive (primitives/Point.cpp:10-26; never hit, pdf 1,
# position = M * p with M already translated by p, Q25).
# ---------------------------------------------------------------------------------------------------------------
def test_sample_one_light_tape_point_emitter(ctx, oracle):
    """uniformSampleOneLight with a point emitter, on GGX surface hits and on a grid medium's boundary hits."""
    for b, center, spread in ((scenes.point_lit_surface_scene(), (0, 1.5, -1), 1.0),
                              (scenes.noise_volume_scene(res=(24, 24, 24), density=12.0, light="point"), (0, 1, -4), 0.8)):
        ctx.upload(b)
        rs = oracle.scene(b)
        o, d = random_rays(3000, 33, center=center, spread=spread)
        d[:, 2] = np.abs(d[:, 2])
        hits = rs.intersect(o, d)
        keep = [i for i in range(len(o)) if hits[i].hit and not hits[i].is_light]
        assert len(keep) > 200
        hh = (abi.Hit * len(keep))(*[hits[i] for i in keep])
        dirs = d[keep]
        seeds = np.arange(len(keep), dtype=np.uint32) + 800
        tape = np.stack([oracle.tape(int(s), 4096) for s in seeds])
        rL, rused = rs.sample_one_light(dirs, hh, seeds)
        L, used = ctx.sample_one_light(dirs, hh, tape)
        same = used == rused
        assert same.mean() > 0.995, f"draw-count agreement {same.mean()}"
        assert_close(L[same], rL[same], what="uniformSampleOneLight(point)", rtol=5e-5, atol=1e-6, frac=0.998)
        assert (rL.sum(axis=1) > 0).mean() > 0.1


def test_li_tape_point_light_surfaces(ctx, oracle):
    b = scenes.point_lit_surface_scene()
    ctx.upload(b)
    rs = oracle.scene(b)
    cp = scenes.CameraParams((0, 2, -5), (0, 1, 0), 45.0)
    cam_o, cam_d = oracle.camera_rays(cp, 1.0, 17, np.random.default_rng(12).uniform(0, 1, (3000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 130000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds)
    same = used == rused
    assert same.mean() > 0.99, f"Li(point, surfaces) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li point-lit surfaces", rtol=1e-4, atol=1e-5, frac=0.995)
    assert (rL.sum(axis=1) > 0).mean() > 0.5


def test_li_tape_point_light_volume(ctx, oracle):
    """BASELINE configs[1]'s shape end to end, draw for draw: heterogeneous grid medium + point emitter."""
    b = scenes.noise_volume_scene(res=(24, 24, 24), density=12.0, light="point")
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(41)
    n = 1500
    o = np.tile([0, 1, -6], (n, 1)).astype(np.float32)
    tgt = np.stack([rng.uniform(-1, 1, n), 1 + rng.uniform(-1, 1, n), np.zeros(n)], 1)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    seeds = np.arange(n, dtype=np.uint32) + 150000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, o, d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"Li(point, volume) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li point-lit volume", rtol=2e-4, atol=1e-6, frac=0.99)
    assert (rL.sum(axis=1) > 0).mean() > 0.05


def test_li_tape_c2_scene(ctx, oracle):
    """The benchmarked scene itself (scenes.c2_scene: FastNoise 256^3, density 100, scale 5, point emitter, camera
    (0,0,-12)): whole Li paths with the reference's global-majorant walk, draw for draw."""
    b = scenes.c2_scene()
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(43)
    xy = np.stack([rng.uniform(0.3, 0.7, 600), rng.uniform(0.15, 0.85, 600)], 1)  # the part of the frame the volume covers
    cam_o, cam_d = oracle.camera_rays(scenes.C2_CAMERA, 1920 / 1080, 19, xy)
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 170000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds, stride=8192)
    assert rused.max() < 8192
    same = used == rused
    assert same.mean() > 0.95, f"Li(C2) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li C2", rtol=5e-4, atol=1e-6, frac=0.99)
    assert (rused > 100).mean() > 0.3


# ---------------------------------------------------------------------------------------------------------------
# Normal-mapped triangles: Triangle.cpp:74-77 + convertNormalFromTextureMap (utils/Math.h:1209-1215), Q24.
# ---------------------------------------------------------------------------------------------------------------
def _normal_map_cases():
    rng = np.random.default_rng(23)
    img = rng.integers(0, 256, (12, 20, 4), dtype=np.uint8)
    img[..., 2] = np.maximum(img[..., 2], 160)  # mostly outward-facing, like a real tangent-space map
    return {"constant": dict(normal_map=(0.35, 0.6, 0.95)), "image": dict(normal_image=img)}


@pytest.mark.parametrize("kind", ["constant", "image"])
def test_intersect_normal_mapped_mesh(ctx, oracle, kind):
    b = scenes.mesh_scene(n=24, **_normal_map_cases()[kind])
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(2)
    n = 20000
    o = np.stack([rng.uniform(-2, 2, n), rng.uniform(0.5, 3, n), rng.uniform(-2, 2, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d[:, 1] = -np.abs(d[:, 1])
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a, r, m = compare_hits(ctx, rs, o, d, what=f"normal-mapped mesh ({kind})", min_frac=0.9995)
    # the map really bends the normals: compare with the same mesh without one
    plain = scenes.mesh_scene(n=24)
    ctx.upload(plain)
    g = hits_to_arrays(ctx.intersect(o, d), n)
    mm = m & (g["hit"] == 1) & (a["inst"] == 0)
    assert mm.sum() > 1000
    assert (np.abs(g["nrm"][mm] - a["nrm"][mm]).max(axis=1) > 1e-3).mean() > 0.9


def test_li_tape_normal_mapped_mesh(ctx, oracle):
    b = scenes.mesh_scene(n=24, **_normal_map_cases()["image"])
    ctx.upload(b)
    rs = oracle.scene(b)
    cam_o, cam_d = oracle.camera_rays(scenes.MESH_CAMERA, 1.0, 27, np.random.default_rng(14).uniform(0, 1, (2000, 2)))
    seeds = np.arange(len(cam_o), dtype=np.uint32) + 190000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, cam_o, cam_d, 6, seeds)
    same = used == rused
    assert same.mean() > 0.98, f"Li(normal-mapped mesh) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what="Li normal-mapped mesh", rtol=2e-4, atol=1e-5, frac=0.99)
    assert rL.mean() > 0.05


# ---------------------------------------------------------------------------------------------------------------
# FAST medium shading (csrc/ne_device.cuh): what production renders run in the phase-function code instead of the
# reference-order IEEE divisions / local frames. Held to the SAME reference vectors and tolerances as the exact code.
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def fast_ctx(ctx):
    ctx.set_fast_shading(True)
    yield ctx
    ctx.set_fast_shading(False)


def test_fast_medium_shading_phase_function(fast_ctx, oracle):
    """VolumeBSDF eval / pdf / sample (HG g = 0, 0.7, -0.3, isotropic) of the FAST versions against the reference: 1e-5
    relative on values, the reference test's tolerance on sampled directions."""
    ctx = fast_ctx
    for phase, g in (("hg", 0.0), ("hg", 0.7), ("hg", -0.3), ("isotropic", 0.0)):
        b = scenes.s2_volume(phase=phase, g=g)
        ctx.upload(b)
        rs = oracle.scene(b)
        rng = np.random.default_rng(6)
        n = 4000
        nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        nrm[:8] = np.eye(3)[[0, 1, 2, 0, 1, 2, 0, 1]] * np.array([1, 1, 1, -1, -1, -1, 1, -1])[:, None]  # axis-aligned normals
        inc = rng.normal(size=(n, 3))
        sc = rng.normal(size=(n, 3)) * 2
        seeds = np.arange(n, dtype=np.uint32) + 7
        tape = np.stack([oracle.tape(int(s), 2) for s in seeds])
        re, rp, rsmp = rs.bsdf(0, inc, sc, nrm, seeds=seeds)
        e, p, smp = ctx.bsdf(0, inc, sc, nrm, tape=tape)
        assert_close(e, re, what=f"fast phase eval {phase} {g}", rtol=1e-5, atol=1e-8)
        assert_close(p, rp, what=f"fast phase pdf {phase} {g}", rtol=1e-5, atol=1e-8)
        assert_close(smp, rsmp, what=f"fast phase sample {phase} {g}", rtol=2e-5, atol=2e-6)
        assert np.abs(np.linalg.norm(smp, axis=1) - 1).max() < 1e-5


def test_fast_medium_shading_rejects_surfaces(fast_ctx):
    fast_ctx.upload(scenes.s1_cornell(with_sphere=False))
    one = np.array([[0.0, 1.0, 0.0]], np.float32)
    with pytest.raises(abi.NarvalB200Error):
        fast_ctx.bsdf(0, one, one, one)


@pytest.mark.parametrize("light", ["point", "rect"])
def test_fast_medium_shading_one_light_tape(fast_ctx, oracle, light):
    """uniformSampleOneLight at medium hits through the FAST versions, draw for draw against the reference."""
    ctx = fast_ctx
    b = scenes.noise_volume_scene(res=(24, 24, 24), density=12.0, light=light)
    ctx.upload(b)
    rs = oracle.scene(b)
    o, d = random_rays(3000, 33, center=(0, 1, -4), spread=0.8)
    d[:, 2] = np.abs(d[:, 2])
    hits = rs.intersect(o, d)
    keep = [i for i in range(len(o)) if hits[i].hit and not hits[i].is_light]
    assert len(keep) > 200
    hh = (abi.Hit * len(keep))(*[hits[i] for i in keep])
    dirs = d[keep]
    seeds = np.arange(len(keep), dtype=np.uint32) + 800
    tape = np.stack([oracle.tape(int(s), 4096) for s in seeds])
    rL, rused = rs.sample_one_light(dirs, hh, seeds)
    L, used = ctx.sample_one_light(dirs, hh, tape)
    same = used == rused
    assert same.mean() > 0.995, f"draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what=f"fast uniformSampleOneLight({light})", rtol=5e-5, atol=1e-6, frac=0.998)
    assert (rL.sum(axis=1) > 0).mean() > 0.1


@pytest.mark.parametrize("light", ["point", "rect"])
def test_fast_medium_shading_li_tape(fast_ctx, oracle, light):
    """Whole Li paths through a heterogeneous medium with FAST shading against the reference, tape for tape: the few
    ulps of difference in a scattered direction must not change what a path does (same draw counts, same radiance)."""
    ctx = fast_ctx
    b = scenes.noise_volume_scene(res=(24, 24, 24), density=12.0, light=light)
    ctx.upload(b)
    rs = oracle.scene(b)
    rng = np.random.default_rng(41)
    n = 1500
    o = np.tile([0, 1, -6], (n, 1)).astype(np.float32)
    tgt = np.stack([rng.uniform(-1, 1, n), 1 + rng.uniform(-1, 1, n), np.zeros(n)], 1)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    seeds = np.arange(n, dtype=np.uint32) + 150000
    L, used, rL, rused = _tape_li(ctx, oracle, rs, o, d, 6, seeds, stride=16384)
    same = used == rused
    assert same.mean() > 0.97, f"fast Li({light}, volume) draw-count agreement {same.mean()}"
    assert_close(L[same], rL[same], what=f"fast Li {light}-lit volume", rtol=2e-4, atol=1e-6, frac=0.99)
    assert (rL.sum(axis=1) > 0).mean() > 0.05
