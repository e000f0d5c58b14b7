"""CPU check of the ALGEBRAIC shortcuts of FAST medium shading (csrc/ne_device.cuh), in numpy float32 with
the reference's order of operations on one side and the shortcut on the other. (The hardware approximations themselves -
reciprocal, rsqrt, sine, cosine at <= 2^-21 relative - are held to the oracle on the GPU: tests/test_gpu_parity.py
test_fast_medium_shading_*.) Reference formulas: core/VolumeBSDF.cpp, materials/Medium.h:73-130 (HG, isotropic),
utils/Math.h:591-631 (generateOrthonormalCS, toLCS), :442-457 (sampleUnitSphere)."""
import numpy as np

f32 = np.float32
INV4PI = f32(0.07957747154594766788)


def normalize(v):
    return (v * (f32(1) / np.sqrt(np.sum(v * v, axis=1, dtype=f32, keepdims=True), dtype=f32))).astype(f32)


def onb(n):
    """generateOrthonormalCS, utils/Math.h:591-599"""
    a = np.abs(n[:, 0]) > np.abs(n[:, 1])
    va = np.stack([-n[:, 2], np.zeros_like(n[:, 0]), n[:, 0]], 1) / np.sqrt(n[:, 0] ** 2 + n[:, 2] ** 2, dtype=f32)[:, None]
    vb = np.stack([np.zeros_like(n[:, 0]), n[:, 2], -n[:, 1]], 1) / np.sqrt(n[:, 1] ** 2 + n[:, 2] ** 2, dtype=f32)[:, None]
    v = np.where(a[:, None], va, vb).astype(f32)
    u = normalize(np.cross(n, v).astype(f32))
    return v, u


def to_lcs(x, n, ss, ts):
    return np.stack([np.sum(x * ss, 1, dtype=f32), np.sum(x * ts, 1, dtype=f32), np.sum(x * n, 1, dtype=f32)], 1).astype(f32)


def hg_reference(g, incoming, scattered, n):
    """VolumeBSDF::eval: wo = toLCS(-normalize(incoming)), wi = toLCS(scattered); HG::eval normalises both again."""
    ss, ts = onb(n)
    wo = to_lcs(-normalize(incoming), n, ss, ts)
    wi = to_lcs(scattered, n, ss, ts)
    c = np.sum(normalize(wo) * normalize(wi), 1, dtype=f32)
    denom = f32(1) + g * g - f32(2) * g * c
    return (INV4PI * (f32(1) - g * g) / (denom * np.sqrt(denom, dtype=f32))).astype(f32)


def hg_shortcut(g, incoming, scattered):
    """phase_eval_fast: the cosine of the two WORLD directions (the local frame is a rotation), one rsqrt of the product."""
    if g == 0:
        return np.full(len(incoming), INV4PI, f32)
    c = -np.sum(incoming * scattered, 1, dtype=f32) / np.sqrt(np.sum(incoming ** 2, 1, dtype=f32) * np.sum(scattered ** 2, 1, dtype=f32), dtype=f32)
    denom = f32(1) + g * g - f32(2) * g * c
    return (INV4PI * (f32(1) - g * g) / denom / np.sqrt(denom, dtype=f32)).astype(f32)


def rays(n, seed):
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=(n, 3)).astype(f32)
    nrm = normalize(nrm)
    return nrm, rng.normal(size=(n, 3)).astype(f32), (rng.normal(size=(n, 3)) * 2).astype(f32)


def test_phase_cosine_does_not_need_the_local_frame():
    nrm, inc, sc = rays(20000, 3)
    for g in (f32(0.7), f32(-0.3), f32(0.05), f32(0.95)):
        ref, fast = hg_reference(g, inc, sc, nrm), hg_shortcut(g, inc, sc)
        rel = np.abs(fast.astype(np.float64) - ref) / ref
        # the frame is orthonormal to a few ulps; HG amplifies an error of the cosine by at most 3 g / (1 - g)^2
        assert rel.max() < 1e-5 * max(1.0, 3 * abs(float(g)) / (1 - abs(float(g))) ** 2 / 20), (g, rel.max())


def test_hg_with_g_zero_is_the_constant_bit_for_bit():
    nrm, inc, sc = rays(20000, 4)
    ref = hg_reference(f32(0), inc, sc, nrm)
    assert np.array_equal(ref, np.full(len(ref), INV4PI, f32))
    assert np.array_equal(hg_shortcut(f32(0), inc, sc), ref)


def test_a_phase_function_divided_by_its_pdf_is_one():
    """volume_collision: fr / pdf of a phase function is x / x - exactly 1 in IEEE arithmetic for finite non-zero x."""
    nrm, inc, sc = rays(20000, 5)
    x = hg_reference(f32(0.6), inc, sc, nrm)
    assert np.array_equal(x / x, np.ones_like(x))


def test_sample_unit_sphere_without_acos():
    """sampleUnitSphere(e1, e2): phi = float(acos(1 - 2 e2)), z = cos(phi), r = sin(phi) in double; the shortcut takes
    z = 1 - 2 e2 and r = sqrt(1 - z^2)."""
    rng = np.random.default_rng(6)
    e1, e2 = rng.uniform(0, 1, 50000).astype(f32), rng.uniform(0, 1, 50000).astype(f32)
    e2[:4] = (0.0, 1.0 - 2.0 ** -24, 0.5, 2.0 ** -24)
    phi = np.arccos((f32(1) - f32(2) * e2).astype(np.float64)).astype(f32)
    theta = (2 * np.pi * e1.astype(np.float64)).astype(f32)
    ref = np.stack([np.sin(phi.astype(np.float64)) * np.cos(theta.astype(np.float64)), np.sin(phi.astype(np.float64)) * np.sin(theta.astype(np.float64)),
                    np.cos(phi.astype(np.float64))], 1).astype(f32)
    z = f32(1) - f32(2) * e2
    r = np.sqrt(np.maximum(f32(0), f32(1) - z * z), dtype=f32)
    turn = f32(2 * np.pi) * (e1 - f32(0.5))  # the kernel evaluates sine / cosine at phi - pi, where the hardware is most accurate
    fast = np.stack([-np.cos(turn, dtype=f32) * r, -np.sin(turn, dtype=f32) * r, z], 1).astype(f32)
    assert np.abs(fast - ref).max() < 2e-6  # the per-function test's absolute tolerance on sampled directions


def test_hg_sample_cosine_with_reciprocals():
    """HG::sample's cos(theta) with the two divisions replaced by multiplications with reciprocals."""
    rng = np.random.default_rng(7)
    u0 = rng.uniform(0, 1, 50000).astype(f32)
    for g in (f32(0.7), f32(-0.3), f32(0.9)):
        sqr = (f32(1) - g * g) / (f32(1) + g - f32(2) * g * u0)
        ref = f32(-1) / (f32(2) * g) * (f32(1) + g * g - sqr * sqr)
        sq2 = (f32(1) - g * g) * (f32(1) / (f32(1) + g - f32(2) * g * u0))
        fast = -(f32(1) / (f32(2) * g)) * (f32(1) + g * g - sq2 * sq2)
        assert np.abs(fast - ref).max() < 2e-6, g
