"""Known-answer tests that pin the CPU oracle (oracle/rtw_oracle.hpp).

The reference has no tests or golden vectors for its render path, so the oracle is
pinned by (a) analytic cases worked out by hand from the cited Rust lines, (b) the one table the
reference carries in a comment (sphere uv, spherical.rs:66-68), (c) published Philox4x32-10 vectors
(Random123 kat_vectors) and (d) a committed golden file of its own outputs (tests/golden).
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

import raytracer_weekend_b200 as rtw
from conftest import bits

INF = np.float32(np.inf)


def one_hit(scene, o, d, time=0.0, t_min=0.001, t_max=np.inf, mode=0):
    return scene.trace_closest(rtw.make_rays([o], [d], time, t_min, t_max), mode)[0]


# ---- RNG -------------------------------------------------------------------------------------------
def test_philox_known_answers(oracle):
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        c = np.array(ctr, np.uint32)
        k = np.array(key, np.uint32)
        out = np.zeros(4, np.uint32)
        oracle.fn("philox4x32_10")(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert tuple(int(x) for x in out) == want


def test_rng_float_distributions(oracle):
    n = 20000
    u = np.zeros(n, np.uint32)
    f = np.zeros(n, np.float32)
    g = np.zeros(n, np.float32)
    oracle.fn("rng_draws")(7, 3, 5, 1, 0, 0.0, 0.0, n, u.ctypes.data, None)
    oracle.fn("rng_draws")(7, 3, 5, 1, 1, 0.0, 0.0, n, None, f.ctypes.data)
    oracle.fn("rng_draws")(7, 3, 5, 1, 2, -1.0, 1.0, n, None, g.ctypes.data)
    # gen::<f32>() = (u32 >> 8) * 2^-24 ; gen_range(-1..1) = v12*2 - 3 with v12 from the top 23 bits
    assert np.array_equal(f, (u >> 8).astype(np.float32) * np.float32(2.0 ** -24))
    v12 = ((u >> 9) | 0x3F800000).view(np.float32)
    assert np.array_equal(g, v12 * np.float32(2.0) + np.float32(-3.0))
    assert f.min() >= 0 and f.max() < 1 and g.min() >= -1 and g.max() < 1
    assert abs(f.mean() - 0.5) < 0.01 and abs(g.mean()) < 0.02
    # streams of different (pixel, sample, stage) differ; same key repeats
    u2 = np.zeros(n, np.uint32)
    oracle.fn("rng_draws")(7, 3, 5, 1, 0, 0.0, 0.0, n, u2.ctypes.data, None)
    assert np.array_equal(u, u2)
    oracle.fn("rng_draws")(7, 3, 5, 2, 0, 0.0, 0.0, n, u2.ctypes.data, None)
    assert not np.array_equal(u, u2)


# ---- spheres ---------------------------------------------------------------------------------------
@pytest.fixture()
def unit_sphere(oracle):
    s = oracle.new_scene()
    m = s.lambertian(s.texture_uvdebug())
    s.sphere((0, 0, 0), 1.0, m)
    s.build()
    yield s
    s.close()


def test_sphere_uv_table(unit_sphere):
    # spherical.rs:66-68
    table = {(1, 0, 0): (0.50, 0.50), (-1, 0, 0): (0.00, 0.50), (0, 1, 0): (0.50, 1.00), (0, -1, 0): (0.50, 0.00),
             (0, 0, 1): (0.25, 0.50), (0, 0, -1): (0.75, 0.50)}
    for p, (u, v) in table.items():
        o = tuple(3.0 * x for x in p)
        d = tuple(-float(x) for x in p)
        h = one_hit(unit_sphere, o, d)
        assert h["prim_id"] == 0 and h["t"] == np.float32(2.0)
        np.testing.assert_allclose(h["p"], p, atol=1e-6)
        np.testing.assert_allclose(h["normal"], p, atol=1e-6)
        assert h["front_face"] == 1
        if p != (-1, 0, 0):  # u wraps 0 == 1 at the seam (atan2(-0, -1) = -pi)
            assert abs(h["u"] - u) < 1e-6
        else:
            assert min(abs(h["u"]), abs(h["u"] - 1.0)) < 1e-6
        assert abs(h["v"] - v) < 1e-6


def test_sphere_roots_and_window(unit_sphere):
    # spherical.rs:31-46: nearest root, else the far root, both limits inclusive
    h = one_hit(unit_sphere, (0, 0, -3), (0, 0, 1))
    assert h["t"] == 2.0 and h["front_face"] == 1
    h = one_hit(unit_sphere, (0, 0, 0), (0, 0, 2))  # from inside: far root, normal flipped
    assert h["t"] == 0.5 and h["front_face"] == 0
    np.testing.assert_allclose(h["normal"], (0, 0, -1))
    assert one_hit(unit_sphere, (0, 0, -3), (0, 0, 1), t_max=1.999)["prim_id"] == -1
    assert one_hit(unit_sphere, (0, 0, -3), (0, 0, 1), t_max=2.0)["prim_id"] == 0   # t == t_max accepted
    assert one_hit(unit_sphere, (0, 0, -3), (0, 0, 1), t_min=2.0)["t"] == 2.0       # t == t_min accepted
    assert one_hit(unit_sphere, (0, 0, -3), (0, 0, 1), t_min=2.0001)["t"] == 4.0    # falls to the far root
    assert one_hit(unit_sphere, (0, 2, -3), (0, 0, 1))["prim_id"] == -1             # miss
    # direction is not normalised: t scales
    assert one_hit(unit_sphere, (0, 0, -3), (0, 0, 4))["t"] == 0.5


def test_negative_radius_sphere(oracle):
    # scenes.rs:90-94: the hollow glass trick; normal = (p - c) / r points inwards
    with oracle.new_scene() as s:
        s.sphere((0, 0, 0), -0.5, s.dielectric(1.5))
        s.build()
        h = one_hit(s, (0, 0, -2), (0, 0, 1))
        assert h["t"] == 1.5
        np.testing.assert_allclose(h["normal"], (0, 0, -1))  # outward (0,0,+1) faces away -> flipped back
        assert h["front_face"] == 0


def test_moving_sphere_center(oracle):
    # spherical.rs:117-123: c(t) = c0 + ((t - t0)/(t1 - t0)) * (c1 - c0)
    with oracle.new_scene() as s:
        s.moving_sphere((0, 0, 0), 0.0, (0, 2, 0), 1.0, 0.5, s.lambertian_rgb(.5, .5, .5))
        s.build()
        assert one_hit(s, (0, 0, -3), (0, 0, 1), time=0.0)["t"] == 2.5
        assert one_hit(s, (0, 0, -3), (0, 0, 1), time=0.5)["prim_id"] == -1
        assert one_hit(s, (0, 1, -3), (0, 0, 1), time=0.5)["t"] == 2.5
        assert one_hit(s, (0, 2, -3), (0, 0, 1), time=1.0)["t"] == 2.5


# ---- rectangles, cuboid, ties ------------------------------------------------------------------------------
def test_rectangles(oracle):
    with oracle.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        assert s.xy_rect(0, 2, 0, 4, 5, m) == 0
        assert s.xz_rect(0, 2, 0, 4, 5, m) == 1
        assert s.yz_rect(0, 2, 0, 4, 5, m) == 2
        s.build()
        h = one_hit(s, (0.5, 1.0, 0), (0, 0, 1))
        assert h["prim_id"] == 0 and h["t"] == 5.0 and h["u"] == 0.25 and h["v"] == 0.25
        np.testing.assert_allclose(h["normal"], (0, 0, -1))
        assert h["front_face"] == 0  # outward normal is +z, the ray travels along +z
        h = one_hit(s, (0.5, 0, 1.0), (0, 1, 0))
        assert h["prim_id"] == 1 and h["t"] == 5.0 and (h["u"], h["v"]) == (0.25, 0.25)
        h = one_hit(s, (0, 0.5, 1.0), (1, 0, 0))
        assert h["prim_id"] == 2 and h["t"] == 5.0 and (h["u"], h["v"]) == (0.25, 0.25)
        # bounds are inclusive (rectangular.rs:40)
        assert one_hit(s, (2.0, 4.0, 0), (0, 0, 1))["prim_id"] == 0
        assert one_hit(s, (2.0001, 4.0, 0), (0, 0, 1))["prim_id"] == -1
        # a ray parallel to the xy plane: t = +-inf fails `t > t_max` (rectangular.rs:34-37) ...
        assert one_hit(s, (0.5, 1.0, 0), (1, 0, 0), t_max=1e30)["prim_id"] != 0
        # ... [QUIRK] but a ray lying IN the plane gives 0/0 = NaN, which fails every comparison of
        # rectangular.rs:35 and :40 and is returned as a hit with t = NaN.
        h = one_hit(s, (0.5, 1.0, 5.0), (1, 0, 0), t_max=1e30)
        assert h["prim_id"] == 0 and np.isnan(h["t"])


def test_cuboid_side_order_and_tie_rule(oracle):
    # rectangular.rs:177-234: XY@z1, XY@z0, XZ@y1, XZ@y0, YZ@x1, YZ@x0 -> ids 0..5
    with oracle.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        assert s.cuboid((0, 0, 0), (1, 1, 1), m) == 0
        assert s.num_prims == 6
        s.build()
        assert one_hit(s, (.5, .5, 3), (0, 0, -1))["prim_id"] == 0
        assert one_hit(s, (.5, .5, -3), (0, 0, 1))["prim_id"] == 1
        assert one_hit(s, (.5, 3, .5), (0, -1, 0))["prim_id"] == 2
        assert one_hit(s, (.5, -3, .5), (0, 1, 0))["prim_id"] == 3
        assert one_hit(s, (3, .5, .5), (-1, 0, 0))["prim_id"] == 4
        assert one_hit(s, (-3, .5, .5), (1, 0, 0))["prim_id"] == 5
        # exact-t tie on the edge x=1,z=1 seen along the diagonal: XY@z1 (id 0) and YZ@x1 (id 4) both
        # report t = 2; "later object wins" (hittable/mod.rs:61-66, t <= closest accepted)
        h = one_hit(s, (3, .5, 3), (-1, 0, -1))
        assert h["t"] == 2.0 and h["prim_id"] == 4
    with oracle.new_scene() as s:  # two coincident rects: the second one wins
        m = s.lambertian_rgb(.5, .5, .5)
        s.xy_rect(0, 1, 0, 1, 2, m)
        s.xy_rect(0, 1, 0, 1, 2, m)
        s.build()
        assert one_hit(s, (.5, .5, 0), (0, 0, 1))["prim_id"] == 1


# ---- triangles ---------------------------------------------------------------------------------------------
def test_triangle(oracle):
    with oracle.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        s.triangles([[0, 0, 0, 2, 0, 0, 0, 2, 0]], m)                       # defaults (triangular.rs:47-65)
        s.triangles([[0, 0, 5, 2, 0, 5, 0, 2, 5]], m, normals=[[0, 0, 1, 0, 0, 1, 0, 0, 1]],
                    uvs=[[.1, .2, .3, .4, .5, .6]])
        s.build()
        h = one_hit(s, (.5, .5, -1), (0, 0, 1))
        assert h["prim_id"] == 0 and h["t"] == 1.0
        # un-normalised face normal (b-a)x(c-a) = (0,0,4); interpolated (1-u-v)n + u n + v n ; faces the ray
        np.testing.assert_allclose(h["normal"], (0, 0, -4))
        assert h["front_face"] == 0
        # default uvs (0,0),(1,0),(0,1): texture uv = barycentric (u, v) = (.25, .25)
        assert (h["u"], h["v"]) == (0.25, 0.25)
        h = one_hit(s, (.5, .5, 6), (0, 0, -1))
        assert h["prim_id"] == 1 and h["t"] == 1.0 and h["front_face"] == 1
        np.testing.assert_allclose(h["normal"], (0, 0, 1))
        np.testing.assert_allclose((h["u"], h["v"]), (.5 * .1 + .25 * .3 + .25 * .5, .5 * .2 + .25 * .4 + .25 * .6), rtol=1e-6)
        # two-sided, edges inclusive, outside rejected, parallel ray (det = 0 -> inf/NaN) rejected
        assert one_hit(s, (0, 0, -1), (0, 0, 1))["prim_id"] == 0
        assert one_hit(s, (1, 1, -1), (0, 0, 1))["prim_id"] == 0       # u + v == 1
        assert one_hit(s, (1.01, 1, -1), (0, 0, 1))["prim_id"] == -1       # u + v = 1.005: outside both
        assert one_hit(s, (1.5, 1.5, -1), (0, 0, 1), t_max=3)["prim_id"] == -1
        assert one_hit(s, (-1, .5, 0), (1, 0, 0), t_max=100)["prim_id"] == -1


# ---- Aabb::hit ---------------------------------------------------------------------------------------------
def test_aabb_hit(oracle):
    def hit(bmin, bmax, o, d, t_min=0.001, t_max=np.inf):
        r = rtw.make_rays([o], [d], 0, t_min, t_max)
        return oracle.fn("aabb_hit")((C.c_float * 3)(*bmin), (C.c_float * 3)(*bmax), r.ctypes.data)

    assert hit((0, 0, 0), (1, 1, 1), (.5, .5, -1), (0, 0, 1)) == 1
    assert hit((0, 0, 0), (1, 1, 1), (.5, .5, 2), (0, 0, -1)) == 1      # negative direction: swap (aabb.rs:33-35)
    assert hit((0, 0, 0), (1, 1, 1), (.5, .5, 2), (0, 0, 1)) == 0       # behind
    assert hit((0, 0, 0), (1, 1, 1), (2, .5, -1), (0, 0, 1)) == 0
    assert hit((0, 0, 0), (1, 1, 1), (.5, .5, -1), (0, 0, 1), t_max=1.0) == 0   # t_max <= t_min rejects (aabb.rs:42)
    assert hit((0, 0, 0), (1, 1, 1), (.5, .5, -1), (0, 0, 1), t_max=1.5) == 1
    # zero direction component on a box face plane: 0 * inf = NaN is ignored by f32::max/min
    assert hit((0, 0, 0), (1, 1, 1), (0.0, .5, -1), (0, 0, 1)) == 1


# ---- transformations -----------------------------------------------------------------------------------------
def test_translation_and_rotation(oracle):
    with oracle.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        s.push_translation((10, 0, 0))
        s.sphere((0, 0, 0), 1.0, m)
        s.pop_transform()
        s.push_translation((0, 0, 20))
        s.push_rotation_y(90.0)
        s.xy_rect(-1, 1, -1, 1, 3, m)   # normal +z in object space; rotated by +90 deg about y -> +x
        s.pop_transform()
        s.pop_transform()
        s.build()
        h = one_hit(s, (10, 0, -5), (0, 0, 1))
        assert h["prim_id"] == 0 and h["t"] == 4.0
        np.testing.assert_allclose(h["p"], (10, 0, -1), atol=1e-6)
        # [QUIRK] transformations.rs:30-37: the inner normal already faces the ray -> front_face true
        assert h["front_face"] == 1
        h = one_hit(s, (10, 0, 0), (0, 0, 1))  # from inside: inner front_face False, wrapper says True
        assert h["t"] == 1.0 and h["front_face"] == 1
        np.testing.assert_allclose(h["normal"], (0, 0, -1), atol=1e-6)
        # rotated rect now sits at x = +3 (plane normal +x), shifted to z in [19, 21]
        h = one_hit(s, (10, 0, 20), (-1, 0, 0), t_max=100)
        assert h["prim_id"] == 1
        np.testing.assert_allclose(h["p"], (3, 0, 20), atol=1e-5)
        assert abs(h["t"] - 7.0) < 1e-5


def test_yrotation_matches_manual_formula(oracle):
    # transformations.rs:122-138, to the bit
    deg = np.float32(15.0)
    rad = deg * np.float32(np.float32(math.pi) / np.float32(180.0))
    s_, c_ = np.float32(math.sin(float(rad))), np.float32(math.cos(float(rad)))
    with oracle.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        s.push_rotation_y(float(deg))
        s.xy_rect(-100, 100, -100, 100, 2, m)
        s.pop_transform()
        s.build()
        o = np.array([0.3, 0.1, -5.0], np.float32)
        d = np.array([0.2, 0.1, 1.0], np.float32)
        h = one_hit(s, o, d)
        oz = s_ * o[0] + c_ * o[2]
        dz = s_ * d[0] + c_ * d[2]
        t = (np.float32(2.0) - oz) / dz
        assert h["t"] == t


# ---- volumes -------------------------------------------------------------------------------------------------------
def test_constant_medium_log10_quirk(oracle):
    """volumes.rs:38-78: t = max(t1, t_min, 0) + (-1/density) * log10(xi) / |d|  — log10, not ln — accepted iff
    it stays inside the boundary; normal (1,0,0), front face, uv (0,0).  xi = the medium's keyed Philox draw."""
    def keyed_xi(prim_id, stage=0):
        c = np.array([prim_id, 0x80000000 | stage, 0, 0], np.uint32)
        k = np.zeros(2, np.uint32)
        out = np.zeros(4, np.uint32)
        oracle.fn("philox4x32_10")(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        return np.float32(out[0] >> 8) * np.float32(2.0 ** -24)

    for density in (0.05, 0.5, 5.0):
        with oracle.new_scene() as s:
            white = s.texture_solid(1, 1, 1)
            m = s.lambertian(white)
            s.xy_rect(-20, 20, -20, 20, 50, m)           # id 0, far behind
            assert s.begin_medium(density, white) == 1
            s.sphere((0, 0, 0), 2.0, m)
            s.end_medium()
            assert s.begin_medium(density, white) == 2
            s.push_translation((10, 0, 0))
            s.cuboid((-1, -1, -1), (1, 1, 1), m)
            s.pop_transform()
            s.end_medium()
            s.build()
            for pid, o, chord in ((1, (0, 0, -5), 4.0), (2, (10, 0, -5), 2.0)):
                d = np.array([0, 0, 2], np.float32)        # |d| = 2: t is in units of d
                h = one_hit(s, o, d)
                xi = keyed_xi(pid)
                hd = np.float32(-1.0 / density) * np.log10(xi)
                t_in = np.float32((5 - chord / 2) / 2)
                if hd <= chord:
                    assert h["prim_id"] == pid
                    np.testing.assert_allclose(h["t"], t_in + hd / np.float32(2), rtol=2e-6)
                    assert tuple(h["normal"]) == (1.0, 0.0, 0.0) and h["front_face"] == 1 and (h["u"], h["v"]) == (0.0, 0.0)
                else:
                    assert h["prim_id"] == 0                # passes through to the wall
                # starting inside: the entry is clamped to t_min (volumes.rs:47-53)
                h = one_hit(s, (o[0], o[1], 0.0), d)
                if hd <= chord / 2:
                    assert h["prim_id"] == pid
                    np.testing.assert_allclose(h["t"], np.float32(0.001) + hd / np.float32(2), rtol=2e-5)


def test_constant_medium_scatter_probability(oracle):
    """P(scatter inside a chord L) = P(xi >= 10^(-density L)) = 1 - 10^(-density L): measure it with the renderer
    (black medium in front of a white background: a scattered path is absorbed)."""
    density, L = 0.3, 2.0
    with oracle.new_scene() as s:
        black = s.texture_solid(0, 0, 0)
        assert s.begin_medium(density, black) == 0
        s.cuboid((-50, -50, 0), (50, 50, L), s.lambertian(black))
        s.end_medium()
        s.build()
        cam = rtw.camera_new((0, 0, -10), (0, 0, 0), (0, 1, 0), 1.0, 1.0)   # narrow view: chords ~ L
        a, st = s.render(cam, s.params(64, 64, 16, seed=3, background=(1, 1, 1)))
        transmitted = float(a.mean()) / 16
        assert abs(transmitted - 10 ** (-density * L)) < 0.01, transmitted


# ---- camera ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("args", [
    ((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0),
    ((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0),
    ((-5, -30, 25), (0, 0, 5), (1, 0, 0), 40.0, 3840 / 2160, 0.0, 10.0),
])
def test_camera_new_host_equals_oracle(oracle, args):
    a = rtw.camera_new(*args)
    b = oracle.camera_new(*args)
    assert bytes(a) == bytes(b)
    # camera.rs:55-56: lens_radius = aperture / 2 ; w = unit(look_from - look_at)
    assert a.lens_radius == np.float32(args[5] / 2)
    w = np.array(args[0], np.float64) - np.array(args[1], np.float64)
    np.testing.assert_allclose(a.w[:], w / np.linalg.norm(w), rtol=1e-6)


# ---- textures ------------------------------------------------------------------------------------------------------
def test_textures(oracle):
    with oracle.new_scene() as s:
        a = s.texture_solid(.2, .3, .1)
        b = s.texture_solid(.9, .9, .9)
        chk = s.texture_checker(a, b, 10.0)  # Checker::new(odd, even, frequency)
        g, p = rtw.perlin_new(5)
        noise = s.texture_noise(g, p, 4.0)
        img = np.zeros((2, 4, 3), np.uint8)
        img[0, 0] = (255, 0, 0)
        img[1, 3] = (0, 0, 255)
        it = s.texture_image(img)
        uvd = s.texture_uvdebug()
        # texture.rs:70-80: sin(f x) sin(f y) sin(f z) < 0 -> odd
        np.testing.assert_allclose(oracle.texture_value(s, chk, 0, 0, (.1, .1, .1)), (.9, .9, .9))
        np.testing.assert_allclose(oracle.texture_value(s, chk, 0, 0, (-.1, .1, .1)), (.2, .3, .1))
        np.testing.assert_allclose(oracle.texture_value(s, uvd, .25, .75, (0, 0, 0)), (.25, .75, 0))
        # image_texture.rs:34-51: v flipped, nearest texel, clamps
        np.testing.assert_allclose(oracle.texture_value(s, it, 0.0, 1.0, (0, 0, 0)), (1, 0, 0))
        np.testing.assert_allclose(oracle.texture_value(s, it, 1.0, 0.0, (0, 0, 0)), (0, 0, 1))
        np.testing.assert_allclose(oracle.texture_value(s, it, 7.0, -3.0, (0, 0, 0)), (0, 0, 1))
        np.testing.assert_allclose(oracle.texture_value(s, it, 0.3, 0.9, (0, 0, 0)), (0, 0, 0))
        # texture.rs:90-94: grey in [0, 1]; perlin.rs turbulence >= 0
        rs = np.random.RandomState(1)
        for _ in range(200):
            c = oracle.texture_value(s, noise, 0, 0, rs.uniform(-50, 50, 3))
            assert c[0] == c[1] == c[2] and 0.0 <= c[0] <= 1.0


def test_perlin_tables():
    g, p = rtw.perlin_new(11)
    np.testing.assert_allclose(np.linalg.norm(g, axis=1), 1.0, rtol=1e-6)   # perlin.rs:17-19
    for a in range(3):
        assert sorted(p[a].tolist()) == list(range(256))                     # a permutation
    # [QUIRK] perlin.rs:43-48: gen_range(0..i) never leaves an element in place at the first swap it takes part in
    assert p[0][255] != 255


# ---- integrator -----------------------------------------------------------------------------------------------------
def test_recursive_vs_iterative_integrator(oracle):
    with rtw.Scene.from_name(oracle, "cornell-box", 1.0) as s:
        p = s.params(32, 32, 16, seed=9)
        it, st_i = oracle.render_ex(s, s.cameras[0], p, integrator=0)
        rc, st_r = oracle.render_ex(s, s.cameras[0], p, integrator=1)
        assert st_i.segments == st_r.segments and st_i.paths == 32 * 32 * 16
        # same terms, different association of the products: a few ulp per path
        np.testing.assert_allclose(it, rc, rtol=2e-5, atol=1e-6)
    with rtw.Scene.from_name(oracle, "jumpy-balls", 16 / 9) as s:
        p = s.params(48, 27, 4, seed=9)
        it, _ = oracle.render_ex(s, s.cameras[0], p, integrator=0)
        rc, _ = oracle.render_ex(s, s.cameras[0], p, integrator=1)
        np.testing.assert_allclose(it, rc, rtol=2e-5, atol=1e-6)


def test_slices_and_partition_are_exact_reorderings(oracle):
    with rtw.Scene.from_name(oracle, "cornell-box", 1.0) as s:
        cam = s.cameras[0]
        full, _ = s.render(cam, s.params(40, 40, 8, seed=2))
        sl, _ = s.render(cam, s.params(40, 40, 8, seed=2, slices=4))
        np.testing.assert_allclose(sl, full, rtol=1e-6)           # only the summation tree differs
        parts = [s.render(cam, s.params(40, 40, 8, seed=2, part_rank=r, part_count=3, tile_size=16))[0] for r in range(3)]
        assert np.array_equal(bits(parts[0] + parts[1] + parts[2]), bits(full))   # x + 0 is exact
        for r in range(3):
            assert np.count_nonzero(parts[r]) > 0
        # sample slices: [0,3) + [3,8) is the frame up to rounding; streams are keyed by absolute sample
        a, _ = s.render(cam, s.params(40, 40, 8, seed=2, sample_begin=0, sample_end=3))
        b, _ = s.render(cam, s.params(40, 40, 8, seed=2, sample_begin=3, sample_end=8))
        np.testing.assert_allclose(a + b, full, rtol=1e-6)


def test_reference_bvh_structure_equals_flat_list(oracle):
    """BvhNode (bvh.rs) over the cow must return the flat list's hit except documented ties / culls."""
    with rtw.Scene.from_name(oracle, "cow-lambert-metal", 16 / 9) as s:
        rays = oracle.capture_rays(s, s.cameras[0], 160, 90, 3, 0, 0)
        flat = s.trace_closest(rays, 0)
        ref = s.trace_closest(rays, 2)
        same = flat["prim_id"] == ref["prim_id"]
        tie = (~same) & (bits(flat["t"]) == bits(ref["t"]))
        assert np.count_nonzero(~same & ~tie) <= 2, "cull disagreements beyond the documented level"
        assert np.count_nonzero(flat["prim_id"] >= 0) > 2000


def test_resolve_rgb8(oracle):
    # main.rs:73-86: sqrt(sum/spp), clamp(0, 0.999) * 255.999 as u8
    with oracle.new_scene() as s:
        acc = np.array([[[0.0, 4.0, 16.0], [1.0, 0.25 * 4, 1e9], [-1.0, np.nan, 3.99]]], np.float32)
        out = s.resolve_rgb8(acc, 4)
        assert out[0, 0].tolist() == [0, 255, 255]
        assert out[0, 1].tolist() == [127, 127, 255]
        assert out[0, 2].tolist() == [0, 0, int(255.999 * min(math.sqrt(3.99 / 4), 0.999))]


# ---- golden file ---------------------------------------------------------------------------------------------------
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_hits_v1.npz")


def test_oracle_matches_committed_golden(oracle):
    """tests/golden/make_golden.py wrote these with the oracle at the commit that introduced it; any later
    change of the oracle's arithmetic shows up here."""
    if not os.path.exists(GOLDEN):
        pytest.skip("golden file not generated yet")
    g = np.load(GOLDEN)
    for scene in ("cornell-box", "cow-lambert-metal", "jumpy-balls"):
        with rtw.Scene.from_name(oracle, scene, 16 / 9 if scene != "cornell-box" else 1.0, seed=1) as s:
            rays = g[f"{scene}/rays"].view(rtw.RAY_DTYPE).reshape(-1)
            hits = s.trace_closest(rays)
            want = g[f"{scene}/hits"].view(rtw.HIT_DTYPE).reshape(-1)
            assert np.array_equal(hits["prim_id"], want["prim_id"])
            assert np.array_equal(bits(hits["t"]), bits(want["t"]))
            assert np.array_equal(bits(hits["normal"]), bits(want["normal"]))


RENDER_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_renders_v1.npz")


def test_oracle_matches_committed_render_golden(oracle):
    """tests/golden/make_golden_renders.py: small frames of seven scenes (every material, texture, wrapper, the media).
    Pins integrator, materials, textures and the Philox stream of the oracle bit for bit."""
    g = np.load(RENDER_GOLDEN)
    scenes = sorted({k.split("/")[0] for k in g.files})
    assert len(scenes) == 7
    for scene in scenes:
        w, h, spp, slices, segments = (int(x) for x in g[f"{scene}/meta"])
        with rtw.Scene.from_name(oracle, scene, w / h, seed=1) as s:
            accum, st = s.render(s.cameras[0], s.params(w, h, spp, seed=4242, slices=slices))
        assert st.segments == segments, scene
        assert np.array_equal(bits(accum), bits(g[f"{scene}/accum"])), scene


def test_oracle_and_kernels_carry_every_constant_of_the_reference_library():
    """tests/golden/scenes_rs_literals.json (tools/make_scene_literals.py) also lists, per source file of raytracer_weekend_lib, the
    floating-point literals other than 0 / 0.5 / 1 / 2: t_min 0.001 (lib.rs), the 0.0001 / 0.0002 box paddings (rectangular.rs,
    triangular.rs), the 0.0001 re-entry offset (volumes.rs), near_zero's 1e-8 (vec3.rs), the Hermite 3.0 (perlin.rs), the checker's
    10.0 (texture.rs), 255.0 (image_texture.rs).  An op-for-op restatement must carry each of them — in the oracle AND in the CUDA
    sources (a constant mistyped in both would pass every parity test)."""
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = json.load(open(os.path.join(root, "tests", "golden", "scenes_rs_literals.json")))["library_literals"]
    want = {v for vals in lib.values() for v in vals}
    assert {0.001, 0.0001, 0.0002, 1e-8, 3.0, 255.0} <= want
    lit = re.compile(r"(?<![\w.])\d+\.\d*(?:e-?\d+)?f?|(?<![\w.])\d+e-?\d+f?")

    def literals(paths):
        got = set()
        for p in paths:
            text = re.sub(r"//[^\n]*", "", open(os.path.join(root, p)).read())
            got |= {float(x.rstrip("f")) for x in lit.findall(text)}
        return got

    oracle = literals(["oracle/rtw_oracle.hpp", "oracle/oracle_capi.cpp", "raytracer-weekend_b200/host/rtw_host.cpp",
                       "raytracer-weekend_b200/host/scenes.cpp", "raytracer-weekend_b200/host/rtw_host.hpp"])
    cuda = literals(["raytracer-weekend_b200/csrc/rtw_device.cuh", "raytracer-weekend_b200/csrc/rtw_bvh.cu",
                     "raytracer-weekend_b200/csrc/rtw_render.cu", "raytracer-weekend_b200/csrc/rtw_traverse.cuh",
                     "raytracer-weekend_b200/csrc/rtw_api.cu", "raytracer-weekend_b200/host/rtw_host.cpp",
                     "raytracer-weekend_b200/host/scenes.cpp", "raytracer-weekend_b200/host/rtw_host.hpp"])
    assert not (want - oracle), ("oracle lacks", want - oracle)
    assert not (want - cuda), ("CUDA path lacks", want - cuda)
