"""A SECOND restatement of the reference's primitive hit() functions, in numpy float32, written from the Rust text alone and compared
BIT FOR BIT with the C++ oracle on random rays.  numpy evaluates every float32 operation separately (no contraction), so agreement
means the two transcriptions order the operations the same way — a slip in one of them (a swapped operand, a different association,
a reciprocal instead of a division) shows up as differing bits.  Not a substitute for running the reference (no Rust toolchain here:
DESIGN.md §2), but independent of the oracle's own golden files."""
import numpy as np
import pytest

import raytracer_weekend_b200 as rtw

F = np.float32


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _dot(a, b):      # vec3.rs:46-48   e0*e0 + e1*e1 + e2*e2, left to right
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _cross(a, b):    # vec3.rs:50-56
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _at(o, d, t):    # ray.rs:25-27   origin + t * direction
    return o + t[:, None] * d


def _face(d, outward):   # hittable/mod.rs:32-47
    front = _dot(d, outward) < 0
    return front, np.where(front[:, None], outward, -outward)


def _rays(n, seed, lo=-3.0, hi=3.0):
    rs = np.random.RandomState(seed)
    o = rs.uniform(lo, hi, (n, 3)).astype(F)
    target = rs.uniform(-1.0, 1.0, (n, 3)).astype(F)
    d = (target - o).astype(F)
    return o, d


def _trace(scene, o, d, t_min=0.001, t_max=np.inf, time=0.0):
    return scene.trace_closest(rtw.make_rays(o, d, time, t_min, t_max))


def test_sphere_hit_bitwise(oracle):
    """spherical.rs:18-60 (hit_sphere), centre / radius arbitrary, rays from outside and inside"""
    c = np.array([0.3, -0.2, 0.1], F)
    r = F(1.25)
    with oracle.new_scene() as s:
        s.sphere(tuple(float(x) for x in c), float(r), s.lambertian_rgb(.5, .5, .5))
        s.build()
        o, d = _rays(20000, 1)
        h = _trace(s, o, d)
    t_min, t_max = F(0.001), F(np.inf)
    oc = o - c
    a = _dot(d, d)
    half_b = _dot(oc, d)
    cc = _dot(oc, oc) - r * r
    with np.errstate(invalid="ignore"):
        disc = half_b * half_b - a * cc
        sq = np.sqrt(np.maximum(disc, F(0)))
        root = (-half_b - sq) / a
        far = (-half_b + sq) / a
    use_far = (root < t_min) | (t_max < root)
    root = np.where(use_far, far, root)
    hit = (disc >= 0) & ~((root < t_min) | (t_max < root))
    assert np.array_equal(h["prim_id"] >= 0, hit) and hit.sum() > 5000 and use_far[hit].sum() > 100
    p = _at(o, d, root)
    outward = (p - c) / r
    front, normal = _face(d, outward)
    assert np.array_equal(_bits(h["t"][hit]), _bits(root[hit]))
    assert np.array_equal(_bits(h["p"][hit]), _bits(p[hit]))
    assert np.array_equal(_bits(h["normal"][hit]), _bits(normal[hit]))
    assert np.array_equal(h["front_face"][hit] != 0, front[hit])
    # get_sphere_uv (spherical.rs:62-78): libm on both sides, so a tolerance instead of bits
    with np.errstate(invalid="ignore"):   # rows that are no hit carry NaN
        theta = np.arccos(-outward[:, 1].astype(np.float64))
        phi = np.arctan2(-outward[:, 2].astype(np.float64), outward[:, 0].astype(np.float64)) + np.pi
    np.testing.assert_allclose(h["u"][hit], (phi / (2 * np.pi))[hit], atol=2e-6)
    np.testing.assert_allclose(h["v"][hit], (theta / np.pi)[hit], atol=2e-6)


def test_triangle_hit_bitwise(oracle):
    """triangular.rs:97-138 with the default normals / uvs of triangular.rs:42-65"""
    va, vb, vc = (np.array(v, F) for v in ([-1.0, -0.5, 0.2], [1.5, -0.25, -0.3], [0.1, 1.75, 0.4]))
    with oracle.new_scene() as s:
        s.triangles([list(map(float, np.concatenate([va, vb, vc])))], s.lambertian_rgb(.5, .5, .5))
        s.build()
        o, d = _rays(20000, 2)
        h = _trace(s, o, d)
    t_min, t_max = F(0.001), F(np.inf)
    ab, ac = vb - va, vc - va
    n = _cross(ab, ac)
    with np.errstate(divide="ignore", invalid="ignore"):
        det = -_dot(d, n[None, :])
        inv = F(1.0) / det
        ao = o - va
        aod = _cross(ao, d)
        u = _dot(ac[None, :], aod) * inv
        v = (-_dot(ab[None, :], aod)) * inv
        t = _dot(ao, n[None, :]) * inv
    hit = ~((t < t_min) | (t > t_max)) & (t >= 0) & (u >= 0) & (v >= 0) & ((u + v) <= 1)
    assert np.array_equal(h["prim_id"] >= 0, hit) and hit.sum() > 2000
    p = _at(o, d, t)
    # interpolate_barycentric over three copies of the face normal (triangular.rs:125-131, 47-55): (1-u-v) n + u n + v n
    w0 = F(1.0) - u - v
    hit_normal = (w0[:, None] * n[None, :] + u[:, None] * n[None, :]) + v[:, None] * n[None, :]
    front, normal = _face(d, hit_normal)
    assert np.array_equal(_bits(h["t"][hit]), _bits(t[hit]))
    assert np.array_equal(_bits(h["p"][hit]), _bits(p[hit]))
    assert np.array_equal(h["front_face"][hit] != 0, front[hit])
    assert np.array_equal(_bits(h["normal"][hit]), _bits(normal[hit]))
    # default uvs (0,0), (1,0), (0,1) through the same interpolation (triangular.rs:57-65, 132)
    zero = np.zeros_like(u)
    tu = (w0 * zero + u * F(1.0)) + v * zero
    tv = (w0 * zero + u * zero) + v * F(1.0)
    assert np.array_equal(_bits(h["u"][hit]), _bits(tu[hit])) and np.array_equal(_bits(h["v"][hit]), _bits(tv[hit]))


@pytest.mark.parametrize("kind", ["xy", "xz", "yz"])
def test_rectangle_hit_bitwise(oracle, kind):
    """rectangular.rs:27-57 / 78-108 / 129-159"""
    a0, a1, b0, b1, k = (F(x) for x in (-0.75, 1.25, -1.5, 0.5, 0.3))
    with oracle.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        getattr(s, f"{kind}_rect")(float(a0), float(a1), float(b0), float(b1), float(k), m)
        s.build()
        o, d = _rays(20000, 3)
        h = _trace(s, o, d)
    ia, ib, ik = {"xy": (0, 1, 2), "xz": (0, 2, 1), "yz": (1, 2, 0)}[kind]
    t_min, t_max = F(0.001), F(np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (k - o[:, ik]) / d[:, ik]
        a = o[:, ia] + t * d[:, ia]
        b = o[:, ib] + t * d[:, ib]
    hit = ~((t < t_min) | (t > t_max)) & ~((a < a0) | (a > a1) | (b < b0) | (b > b1))
    assert np.array_equal(h["prim_id"] >= 0, hit) and hit.sum() > 1000
    p = _at(o, d, t)
    outward = np.zeros((len(o), 3), F)
    outward[:, ik] = 1.0
    front, normal = _face(d, outward)
    assert np.array_equal(_bits(h["t"][hit]), _bits(t[hit]))
    assert np.array_equal(_bits(h["p"][hit]), _bits(p[hit]))
    assert np.array_equal(_bits(h["normal"][hit]), _bits(normal[hit]))
    assert np.array_equal(h["front_face"][hit] != 0, front[hit])
    assert np.array_equal(_bits(h["u"][hit]), _bits(((a - a0) / (a1 - a0))[hit]))
    assert np.array_equal(_bits(h["v"][hit]), _bits(((b - b0) / (b1 - b0))[hit]))


def test_aabb_hit_matches(oracle):
    """aabb.rs:23-48: reciprocal per axis, swap on a negative reciprocal, f32::max / min (which drop NaN), reject on t_max <= t_min"""
    import ctypes as C
    rs = np.random.RandomState(4)
    n = 4000
    lo = rs.uniform(-1.0, 0.5, (n, 3)).astype(F)
    hi = (lo + rs.uniform(0.0, 1.5, (n, 3)).astype(F)).astype(F)
    o, d = _rays(n, 5)
    d[rs.rand(n) < 0.1, 0] = 0.0                      # rays parallel to an axis: 1/0 = inf, 0 * inf = NaN paths
    on_face = rs.rand(n) < 0.05
    o[on_face, 0] = lo[on_face, 0]
    rays = rtw.make_rays(o, d, 0.0, 0.001, np.inf)
    got = np.array([oracle.fn("aabb_hit")((C.c_float * 3)(*map(float, lo[i])), (C.c_float * 3)(*map(float, hi[i])),
                                          rays[i:i + 1].ctypes.data) for i in range(n)], bool)
    t_min = np.full(n, 0.001, F)
    t_max = np.full(n, np.inf, F)
    alive = np.ones(n, bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            inv = F(1.0) / d[:, a]
            t0 = (lo[:, a] - o[:, a]) * inv
            t1 = (hi[:, a] - o[:, a]) * inv
            swap = inv < 0
            t0, t1 = np.where(swap, t1, t0), np.where(swap, t0, t1)
            t_min = np.where(np.isnan(t0), t_min, np.where(np.isnan(t_min), t0, np.maximum(t0, t_min)))   # f32::max
            t_max = np.where(np.isnan(t1), t_max, np.where(np.isnan(t_max), t1, np.minimum(t1, t_max)))   # f32::min
            alive &= ~(t_max <= t_min)
    assert np.array_equal(got, alive) and 500 < alive.sum() < n - 500


@pytest.mark.parametrize("args", [
    ((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0),
    ((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0),
    ((-5, -30, 25), (0, 0, 5), (1, 0, 0), 40.0, 3840 / 2160, 0.0, 10.0),
    ((478, 278, -600), (278, 278, 278), (0, 1, 0), 40.0, 1.0, 1.0, 1077.6),
])
def test_camera_new_against_numpy(oracle, args):
    """camera.rs:25-64.  The frame (w, u, v), origin and lens radius use only IEEE-exact operations: bit for bit.  The viewport
    goes through tan(): 1e-6 relative."""
    look_from, look_at, up, vfov, aspect, aperture, focus = args
    cam = oracle.camera_new(*args)
    lf, la, upv = (np.array(x, F) for x in (look_from, look_at, up))

    def unit(a):   # vec3.rs:85-87: a / length(a)
        return a / np.sqrt(_dot(a, a))

    w = unit(lf - la)
    u = unit(_cross(upv, w))
    v = _cross(w, u)
    assert np.array_equal(_bits(np.array(cam.w[:], F)), _bits(w))
    assert np.array_equal(_bits(np.array(cam.u[:], F)), _bits(u))
    assert np.array_equal(_bits(np.array(cam.v[:], F)), _bits(v))
    assert np.array_equal(_bits(np.array(cam.origin[:], F)), _bits(lf))
    assert F(cam.lens_radius) == F(aperture) / F(2.0)
    theta = F(vfov) * (F(np.pi) / F(180.0))               # f32::to_radians
    h = F(np.tan(np.float64(theta / F(2.0))))
    vh = F(2.0) * h
    vw = F(aspect) * vh
    horizontal = (F(focus) * vw) * u
    vertical = (F(focus) * vh) * v
    llc = ((lf - horizontal / F(2.0)) - vertical / F(2.0)) - F(focus) * w
    np.testing.assert_allclose(cam.horizontal[:], horizontal, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(cam.vertical[:], vertical, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(cam.lower_left_corner[:], llc, rtol=2e-6, atol=2e-4)


# ---- Material::scatter (material.rs:40-147) from the path stream ---------------------------------------------------------------
# The random STREAM is this repo's specification (Philox keyed by pixel / sample, stage = bounce + 1: DESIGN.md §2), so the draws
# are taken from the oracle's own generator; what is re-derived here from the Rust text is everything done WITH them.
W, H, SEED = 48, 27, 77


def _draws(oracle, pixel, stage, kind, n, lo=-1.0, hi=1.0):
    import ctypes as C
    u = np.zeros(n, np.uint32)
    f = np.zeros(n, F)
    oracle.fn("rng_draws")(SEED, int(pixel), 0, int(stage), kind, lo, hi, n, u.ctypes.data, f.ctypes.data)
    return f


def _unit(a):
    return a / np.sqrt(_dot(a, a))


def _in_unit_sphere(g):   # vec3.rs:101-108 over the sequential gen_range(-1, 1) draws g
    for k in range(len(g) // 3):
        p = g[3 * k:3 * k + 3]
        if _dot(p, p) < F(1.0):
            return p, 3 * (k + 1)
    raise AssertionError("not enough draws")


def _reflect(v, n):       # vec3.rs:140-142   v - (2 * dot(v, n)) * n
    return v - (F(2.0) * _dot(v, n)) * n


def _scatter_scene(oracle, material):
    s = oracle.new_scene()
    m = material(s)
    s.sphere((0.0, 0.0, 0.0), 1.5, m)
    s.build()
    cam = oracle.camera_new((0.5, 0.8, -6.0), (0, 0, 0), (0, 1, 0), 35.0, W / H, 0.0, 6.0)
    r0 = oracle.capture_rays(s, cam, W, H, SEED, 0, 0)
    r1 = oracle.capture_rays(s, cam, W, H, SEED, 0, 1)
    hits = s.trace_closest(r0)
    return s, r0, r1, hits


def _pixel_of(idx):       # the reference's pixel order: rows top-down in the buffer, Pixel.row bottom-up (lib.rs:58)
    y_top, col = divmod(idx, W)
    return (H - 1 - y_top) * W + col


def test_lambertian_scatter_bitwise(oracle):
    """material.rs:40-56: normal + random_unit_vector, the near-zero guard, origin = hit point, time kept"""
    s, r0, r1, hits = _scatter_scene(oracle, lambda s: s.lambertian_rgb(.5, .5, .5))
    checked = 0
    for i in np.nonzero(hits["prim_id"] >= 0)[0]:
        g = _draws(oracle, _pixel_of(i), 1, 2, 90)
        p, _ = _in_unit_sphere(g)
        d = hits["normal"][i] + _unit(p)
        if (np.abs(d) < F(1e-8)).all():
            d = hits["normal"][i]
        assert r1["t_max"][i] > r1["t_min"][i]
        assert np.array_equal(_bits(r1["origin"][i]), _bits(hits["p"][i]))
        assert np.array_equal(_bits(r1["direction"][i]), _bits(d)), i
        assert r1["time"][i] == r0["time"][i]
        checked += 1
    assert checked > 200
    assert (r1["t_max"][hits["prim_id"] < 0] < r1["t_min"][hits["prim_id"] < 0]).all()   # a miss ends the path
    s.close()


@pytest.mark.parametrize("fuzz", [0.0, 0.4, 1.0])
def test_metal_scatter_bitwise(oracle, fuzz):
    """material.rs:77-98: reflect(unit(d), n) + fuzz * random_in_unit_sphere (drawn even for fuzz 0); absorbed unless dot(out, n) > 0"""
    s, r0, r1, hits = _scatter_scene(oracle, lambda s: s.metal(.8, .8, .9, fuzz))
    alive = absorbed = 0
    for i in np.nonzero(hits["prim_id"] >= 0)[0]:
        n = hits["normal"][i]
        g = _draws(oracle, _pixel_of(i), 1, 2, 90)
        p, _ = _in_unit_sphere(g)
        out = _reflect(_unit(r0["direction"][i]), n) + F(fuzz) * p
        if _dot(out, n) > 0:
            assert r1["t_max"][i] > r1["t_min"][i]
            assert np.array_equal(_bits(r1["direction"][i]), _bits(out)), i
            assert np.array_equal(_bits(r1["origin"][i]), _bits(hits["p"][i]))
            alive += 1
        else:
            assert r1["t_max"][i] < r1["t_min"][i]
            absorbed += 1
    assert alive > 200 and (fuzz < 1.0 or absorbed > 0)
    s.close()


def test_dielectric_scatter_bitwise(oracle):
    """material.rs:100-147 + vec3.rs:144-151: refraction ratio by face, Schlick with powi(5) = x * (x^2)^2, the draw only when
    refraction is possible (short circuit), reflect / refract"""
    ir = F(1.5)
    s, r0, r1, hits = _scatter_scene(oracle, lambda s: s.dielectric(float(ir)))
    reflected = refracted = 0
    for i in np.nonzero(hits["prim_id"] >= 0)[0]:
        n = hits["normal"][i]
        ratio = F(1.0) / ir if hits["front_face"][i] else ir
        ud = _unit(r0["direction"][i])
        cos_t = np.minimum(_dot(-ud, n), F(1.0))
        sin_t = np.sqrt(F(1.0) - cos_t * cos_t)
        cannot = ratio * sin_t > F(1.0)
        if not cannot:
            r0_ = (F(1.0) - ratio) / (F(1.0) + ratio)
            r0_ = r0_ * r0_
            x = F(1.0) - cos_t
            x2 = x * x
            refl = r0_ + (F(1.0) - r0_) * (x * (x2 * x2))      # powi(5): compiler-rt's square-and-multiply
            xi = _draws(oracle, _pixel_of(i), 1, 1, 1)[0]
            cannot = refl > xi
        if cannot:
            out = _reflect(ud, n)
            reflected += 1
        else:
            cos2 = np.minimum(_dot(-ud, n), F(1.0))
            perp = ratio * (ud + cos2 * n)
            par = (-np.sqrt(np.abs(F(1.0) - _dot(perp, perp)))) * n
            out = perp + par
            refracted += 1
        assert r1["t_max"][i] > r1["t_min"][i]
        assert np.array_equal(_bits(r1["direction"][i]), _bits(out)), (i, bool(cannot))
        assert np.array_equal(_bits(r1["origin"][i]), _bits(hits["p"][i]))
    assert refracted > 150 and reflected > 5
    s.close()


# ---- the whole path: lib.rs:78-117 (sample_pixel, sample_ray), camera.rs:66-74 (get_ray), the list rule of hittable/mod.rs:57-69 --
# A tiny scene path-traced by a second, scalar numpy-float32 implementation written from the Rust text; the frame must equal the
# oracle's RECURSIVE integrator bit for bit (same operations in the same order) — and the oracle's iterative form, which is what the
# GPU is compared with, within the bound test_recursive_vs_iterative_integrator states.
class _Stream:
    """The repo's random stream: raw Philox words from the oracle's generator, mapped like rand's f32 distributions."""

    def __init__(self, oracle, seed, pixel, sample, stage):
        import ctypes as C
        self.words = np.zeros(256, np.uint32)
        dummy = np.zeros(256, F)
        oracle.fn("rng_draws")(seed, int(pixel), int(sample), int(stage), 0, 0.0, 1.0, 256, self.words.ctypes.data, dummy.ctypes.data)
        self.i = 0

    def _next(self):
        w = self.words[self.i]
        self.i += 1
        return w

    def gen(self):                                   # rand Standard: 24 bits
        return F(int(self._next()) >> 8) * F(1.0 / 16777216.0)

    def gen_range(self, lo, hi):                     # rand UniformFloat::sample_single
        v12 = np.array([0x3F800000 | (int(self._next()) >> 9)], np.uint32).view(F)[0]
        scale = F(hi) - F(lo)
        return v12 * scale + (F(lo) - scale)


def _v(*x):
    return np.array(x, F)


def _sphere_hit(o, d, c, r, t_min, t_max):
    oc = o - c
    a = _dot(d, d)
    half_b = _dot(oc, d)
    cc = _dot(oc, oc) - r * r
    disc = half_b * half_b - a * cc
    if disc < 0:
        return None
    sq = np.sqrt(disc)
    root = (-half_b - sq) / a
    if root < t_min or t_max < root:
        root = (-half_b + sq) / a
        if root < t_min or t_max < root:
            return None
    p = o + root * d
    outward = (p - c) / r
    front = _dot(d, outward) < 0
    return root, p, (outward if front else -outward), front


def _xz_rect_hit(o, d, x0, x1, z0, z1, k, t_min, t_max):
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (k - o[1]) / d[1]
    if t < t_min or t > t_max:
        return None
    x = o[0] + t * d[0]
    z = o[2] + t * d[2]
    if x < x0 or x > x1 or z < z0 or z > z1:
        return None
    outward = _v(0, 1, 0)
    front = _dot(d, outward) < 0
    return t, o + t * d, (outward if front else -outward), front


def test_tiny_path_tracer_equals_the_oracle_frame(oracle):
    w, h, spp, seed = 12, 8, 3, 4242
    background = _v(0.3, 0.4, 0.6)
    ground = dict(kind="sphere", c=_v(0, -100.5, -1), r=F(100.0), mat=("lambert", _v(0.8, 0.7, 0.2)))
    ball = dict(kind="sphere", c=_v(0.3, 0.0, -1.2), r=F(0.5), mat=("metal", _v(0.8, 0.8, 0.9), F(0.3)))
    lamp = dict(kind="xz", x0=F(-1.5), x1=F(1.5), z0=F(-2.5), z1=F(0.5), k=F(1.6), mat=("light", _v(3.0, 2.5, 2.0)))
    glass = dict(kind="sphere", c=_v(-0.7, -0.1, -0.6), r=F(0.4), mat=("glass", F(1.5)))
    objects = [ground, ball, lamp, glass]
    with oracle.new_scene() as s:
        s.sphere(tuple(map(float, ground["c"])), float(ground["r"]), s.lambertian_rgb(0.8, 0.7, 0.2))
        s.sphere(tuple(map(float, ball["c"])), float(ball["r"]), s.metal(0.8, 0.8, 0.9, 0.3))
        s.xz_rect(-1.5, 1.5, -2.5, 0.5, 1.6, s.diffuse_light(s.texture_solid(3.0, 2.5, 2.0)))
        s.sphere(tuple(map(float, glass["c"])), float(glass["r"]), s.dielectric(1.5))
        s.build()
        cam = oracle.camera_new((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 50.0, w / h, 0.15, 3.4)
        p = s.params(w, h, spp, seed=seed, slices=1, background=tuple(map(float, background)))
        frame_rec, _ = oracle.render_ex(s, cam, p, mode=0, integrator=1)
        frame_it, _ = oracle.render_ex(s, cam, p, mode=0, integrator=0)
    origin, llc, hor, ver = (np.array(x[:], F) for x in (cam.origin, cam.lower_left_corner, cam.horizontal, cam.vertical))
    cu, cv, lens = np.array(cam.u[:], F), np.array(cam.v[:], F), F(cam.lens_radius)

    def world_hit(o, d, t_min, t_max):               # hittable/mod.rs:57-69
        best, closest = None, t_max
        for ob in objects:
            if ob["kind"] == "sphere":
                r = _sphere_hit(o, d, ob["c"], ob["r"], t_min, closest)
            else:
                r = _xz_rect_hit(o, d, ob["x0"], ob["x1"], ob["z0"], ob["z1"], ob["k"], t_min, closest)
            if r is not None:
                closest = r[0]
                best = (r, ob)
        return best

    def sample_ray(o, d, pixel, sample, depth):      # lib.rs:97-117, recursive as written
        if depth == 0:
            return _v(0, 0, 0)
        hit = world_hit(o, d, F(0.001), F(np.inf))
        if hit is None:
            return background
        (t, pnt, normal, front), ob = hit
        rng = _Stream(oracle, seed, pixel, sample, 50 - depth + 1)
        mat = ob["mat"]
        if mat[0] == "light":                        # light_source.rs: emits, never scatters
            return mat[1]
        if mat[0] == "lambert":
            while True:
                q = _v(rng.gen_range(-1, 1), rng.gen_range(-1, 1), rng.gen_range(-1, 1))
                if _dot(q, q) < 1:
                    break
            out = normal + _unit(q)
            if (np.abs(out) < F(1e-8)).all():
                out = normal
            att = mat[1]
        elif mat[0] == "glass":                      # material.rs:100-147
            ratio = F(1.0) / mat[1] if front else mat[1]
            ud = _unit(d)
            cos_t = np.minimum(_dot(-ud, normal), F(1.0))
            sin_t = np.sqrt(F(1.0) - cos_t * cos_t)
            reflect_it = ratio * sin_t > F(1.0)
            if not reflect_it:
                r0_ = (F(1.0) - ratio) / (F(1.0) + ratio)
                r0_ = r0_ * r0_
                x = F(1.0) - cos_t
                x2 = x * x
                reflect_it = r0_ + (F(1.0) - r0_) * (x * (x2 * x2)) > rng.gen()
            if reflect_it:
                out = _reflect(ud, normal)
            else:
                perp = ratio * (ud + np.minimum(_dot(-ud, normal), F(1.0)) * normal)
                out = perp + (-np.sqrt(np.abs(F(1.0) - _dot(perp, perp)))) * normal
            att = _v(1, 1, 1)
        else:
            refl = _reflect(_unit(d), normal)
            while True:
                q = _v(rng.gen_range(-1, 1), rng.gen_range(-1, 1), rng.gen_range(-1, 1))
                if _dot(q, q) < 1:
                    break
            out = refl + mat[2] * q
            if not _dot(out, normal) > 0:
                return _v(0, 0, 0)                   # emitted = black
            att = mat[1]
        return _v(0, 0, 0) + att * sample_ray(pnt, out, pixel, sample, depth - 1)

    frame = np.zeros((h, w, 3), F)
    for y_top in range(h):
        row = h - 1 - y_top
        for col in range(w):
            pixel = row * w + col
            color = _v(0, 0, 0)
            for smp in range(spp):                   # lib.rs:83-88
                rng = _Stream(oracle, seed, pixel, smp, 0)
                u = (F(col) + rng.gen()) / F(w - 1)
                v = (F(row) + rng.gen()) / F(h - 1)
                while True:                          # camera.rs:66-74, vec3.rs:124-131
                    rd = _v(rng.gen_range(-1, 1), rng.gen_range(-1, 1), 0)
                    if _dot(rd, rd) < 1:
                        break
                rd = lens * rd
                offset = cu * rd[0] + cv * rd[1]
                time = rng.gen_range(float(cam.time0), float(cam.time1))
                o = origin + offset
                d = (((llc + u * hor) + v * ver) - origin) - offset
                color = color + sample_ray(o, d, pixel, smp, 50)
            frame[y_top, col] = color
    assert np.array_equal(_bits(frame), _bits(frame_rec)), "the recursive integrator differs from the Rust text"
    np.testing.assert_allclose(frame_it, frame_rec, rtol=1e-4, atol=1e-5)     # iterative form: same terms, another association
    assert len(np.unique(_bits(frame).reshape(-1, 3), axis=0)) > 12            # a real picture (sky, lamp, ground, metal, multi-bounce mixes)


def test_translation_and_rotation_wrappers_bitwise(oracle):
    """transformations.rs:22-37, 113-147: Translation(YRotation(sphere)) as in scenes.rs' Cornell boxes — the ray is moved and rotated
    into the object's space, the hit point and normal come back out, and BOTH wrappers re-run new_with_face_normal on the normal the
    inner hit already flipped (so a hit through a wrapper always reports front_face = true: the reference's quirk, kept)."""
    c, r = _v(0.2, -0.1, 0.3), F(0.9)
    off = _v(0.7, -0.4, 1.1)
    sin_t, cos_t = F(np.sin(np.float64(0.4))), F(np.cos(np.float64(0.4)))
    with oracle.new_scene() as s:
        s.push_translation(tuple(map(float, off)))
        s.push_rotation_y_sincos(float(sin_t), float(cos_t))
        s.sphere(tuple(map(float, c)), float(r), s.lambertian_rgb(.5, .5, .5))
        s.pop_transform()
        s.pop_transform()
        s.build()
        o, d = _rays(6000, 9)
        h = _trace(s, o, d)
    n_hit = 0
    for i in range(len(o)):
        o1 = o[i] - off                                              # Translation::hit
        o2 = _v(cos_t * o1[0] - sin_t * o1[2], o1[1], sin_t * o1[0] + cos_t * o1[2])   # YRotation::hit
        d2 = _v(cos_t * d[i][0] - sin_t * d[i][2], d[i][1], sin_t * d[i][0] + cos_t * d[i][2])
        got = _sphere_hit(o2, d2, c, r, F(0.001), F(np.inf))
        if got is None:
            assert h["prim_id"][i] < 0
            continue
        t, p, normal, _front = got
        p_r = _v(cos_t * p[0] + sin_t * p[2], p[1], -sin_t * p[0] + cos_t * p[2])
        n_r = _v(cos_t * normal[0] + sin_t * normal[2], normal[1], -sin_t * normal[0] + cos_t * normal[2])
        front_r = _dot(d2, n_r) < 0                                  # against the ROTATED ray (transformations.rs:137-146)
        n_r = n_r if front_r else -n_r
        p_t = p_r + off
        front_t = _dot(d[i], n_r) < 0                                # Translation: against the translated ray (same direction)
        n_t = n_r if front_t else -n_r
        assert h["prim_id"][i] == 0 and _bits(h["t"][i:i + 1])[0] == _bits(np.array([t], F))[0]
        assert np.array_equal(_bits(h["p"][i]), _bits(p_t)), i
        assert np.array_equal(_bits(h["normal"][i]), _bits(n_t)), i
        assert bool(h["front_face"][i]) == bool(front_t)
        n_hit += 1
    assert n_hit > 500


def test_moving_sphere_bitwise(oracle):
    """spherical.rs:117-123 + hit_sphere: centre(time) = c0 + ((time - t0) / (t1 - t0)) * (c1 - c0)"""
    c0, c1, t0, t1, r = _v(-0.3, 0.0, 0.2), _v(0.4, 0.6, 0.1), F(0.25), F(1.5), F(0.8)
    with oracle.new_scene() as s:
        s.moving_sphere(tuple(map(float, c0)), float(t0), tuple(map(float, c1)), float(t1), float(r), s.lambertian_rgb(.5, .5, .5))
        s.build()
        o, d = _rays(4000, 11)
        times = np.random.RandomState(12).uniform(0.0, 2.0, len(o)).astype(F)
        rays = rtw.make_rays(o, d, 0.0, 0.001, np.inf)
        rays["time"] = times
        h = s.trace_closest(rays)
    n_hit = 0
    for i in range(len(o)):
        centre = c0 + ((times[i] - t0) / (t1 - t0)) * (c1 - c0)
        got = _sphere_hit(o[i], d[i], centre, r, F(0.001), F(np.inf))
        assert (got is None) == (h["prim_id"][i] < 0)
        if got is not None:
            assert _bits(h["t"][i:i + 1])[0] == _bits(np.array([got[0]], F))[0]
            assert np.array_equal(_bits(h["p"][i]), _bits(got[1])) and np.array_equal(_bits(h["normal"][i]), _bits(got[2]))
            n_hit += 1
    assert n_hit > 400


def test_tonemap_against_numpy(oracle):
    """console_app/src/main.rs:73-86: sqrt(sum / spp) clamped to [0, 0.999], * 255.999 as u8"""
    import ctypes as C
    rs = np.random.RandomState(3)
    acc = np.concatenate([rs.uniform(0, 40, 3000), [0.0, -1.0, 1e9, np.nan, 16.0, 15.99999]]).astype(F)
    n = len(acc) // 3 * 3
    acc = acc[:n]
    spp = 16
    out = np.zeros(n, np.uint8)
    oracle.fn("resolve_rgb8").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    oracle.fn("resolve_rgb8")(None, acc.ctypes.data, n // 3, 1, spp, out.ctypes.data)
    with np.errstate(invalid="ignore"):
        c = np.sqrt((F(1.0) / F(spp)) * acc)
        cl = np.where(c < 0, F(0), np.where(c > F(0.999), F(0.999), c))
        v = F(255.999) * cl
    want = np.where(np.isnan(v), 0, np.clip(np.nan_to_num(v), 0, 255)).astype(np.uint8)
    assert np.array_equal(out, want)


# ---- textures: perlin.rs, texture.rs, image_texture.rs ----------------------------------------------------------------------------
def _perlin_noise(grad, perm, p):          # perlin.rs:50-75, 91-123
    fl = np.floor(p)
    base = fl.astype(np.int64)             # to_i64().to_usize(): the wrap of a negative index is undone by & 255
    pw = p - fl
    pf = (pw * pw) * (_v(3, 3, 3) - F(2.0) * pw)       # filter_hermit
    accum = F(0.0)
    one = _v(1, 1, 1)
    for a in range(2):
        for r in range(2):
            for c in range(2):
                idx = (base + np.array([a, r, c])) & 255
                g = grad[perm[0][idx[0]] ^ perm[1][idx[1]] ^ perm[2][idx[2]]]
                cur = _v(a, r, c)
                weight_v = pf - cur        # (the shadowed, FILTERED point: perlin.rs:92,103)
                blend = cur * pf + (one - cur) * (one - pf)
                accum = accum + ((blend[0] * blend[1]) * blend[2]) * _dot(g, weight_v)
    return accum


def _turbulence(grad, perm, p, depth=7):   # perlin.rs:77-89
    accum, temp, weight = F(0.0), p.copy(), F(1.0)
    for _ in range(depth):
        accum = accum + weight * _perlin_noise(grad, perm, temp)
        weight = weight * F(0.5)
        temp = temp * F(2.0)
    return np.abs(accum)


def test_noise_texture_against_numpy(oracle):
    """texture.rs:83-95 over perlin.rs: 0.5 * (1 + sin(scale * z + 10 * turbulence(p, 7))).  Everything inside the sine is IEEE-exact
    arithmetic; the sine itself is libm on both sides (correctly rounded almost always): >= 98 % of the samples must agree bit for
    bit and the rest within 2 ulp — an operation out of order inside noise() / turbulence() would break nearly all of them."""
    grad, perm = rtw.perlin_new(7)
    scale = F(4.0)
    rs = np.random.RandomState(21)
    pts = rs.uniform(-6.0, 6.0, (300, 3)).astype(F)
    with oracle.new_scene() as s:
        tex = s.texture_noise(grad, perm, float(scale))
        exact = 0
        for p in pts:
            got = oracle.texture_value(s, tex, 0.0, 0.0, p)
            x = scale * p[2] + F(10.0) * _turbulence(grad, perm, p)
            want = F(0.5) * (F(1.0) + F(np.sin(np.float64(x))))
            assert got[0] == got[1] == got[2]
            ulp = abs(int(_bits(got[:1])[0]) - int(_bits(np.array([want], F))[0]))
            assert ulp <= 2, (p, got[0], want)
            exact += ulp == 0
    assert exact >= 0.98 * len(pts), exact


def test_checker_uvdebug_and_image_textures_against_numpy(oracle):
    """texture.rs:62-81 (sign of the product of three sines picks odd / even), :97-104 (uv as colour), image_texture.rs:34-51 (clamp, flip v,
    truncate, clamp the index, / 255 as a multiplication by 1/255)"""
    rs = np.random.RandomState(5)
    img = rs.randint(0, 256, (7, 5, 3)).astype(np.uint8)      # height 7, width 5
    with oracle.new_scene() as s:
        odd, even = s.texture_solid(.2, .3, .1), s.texture_solid(.9, .9, .9)
        chk = s.texture_checker(odd, even, 10.0)
        uvd = s.texture_uvdebug()
        imt = s.texture_image(img)
        for _ in range(300):
            p = rs.uniform(-3, 3, 3).astype(F)
            u, v = (F(x) for x in rs.uniform(-0.2, 1.2, 2))
            sines = np.sin(np.float64(F(10.0) * p[0])) * np.sin(np.float64(F(10.0) * p[1])) * np.sin(np.float64(F(10.0) * p[2]))
            if abs(sines) > 1e-5:                                 # away from the sign change, where libm's last bit cannot matter
                want = (.2, .3, .1) if sines < 0 else (.9, .9, .9)
                np.testing.assert_array_equal(oracle.texture_value(s, chk, float(u), float(v), p), np.array(want, F))
            np.testing.assert_array_equal(oracle.texture_value(s, uvd, float(u), float(v), p), _v(u, v, 0))
            uc = np.clip(u, F(0), F(1))
            vc = F(1.0) - np.clip(v, F(0), F(1))
            i = min(max(int(uc * F(5)), 0), 4)
            j = min(max(int(vc * F(7)), 0), 6)
            want = img[j, i].astype(F) * (F(1.0) / F(255.0))
            assert np.array_equal(_bits(oracle.texture_value(s, imt, float(u), float(v), p)), _bits(want))


def test_constant_medium_against_numpy(oracle):
    """volumes.rs:38-78 over a sphere boundary: two boundary queries ((-inf, inf), then (t1 + 0.0001, inf)), clamp to the window, max(0),
    hit_distance = (-1 / density) * log10(xi) — log10, the reference's quirk — and t = t1 + hit_distance / |d|; normal (1, 0, 0), front
    face true.  xi is this repo's keyed draw: word 0 of Philox block (medium id, 0x80000000 | stage) under the path's key;
    a bare ray batch has the all-zero key."""
    import ctypes as C
    c, r, density = _v(0.1, 0.2, -0.3), F(1.4), F(0.9)
    ctr = (C.c_uint32 * 4)(0, 0x80000000, 0, 0)      # medium id 0, stage 0, seed 0
    key = (C.c_uint32 * 2)(0, 0)
    out = (C.c_uint32 * 4)()
    oracle.fn("philox4x32_10")(ctr, key, out)
    xi = F(out[0] >> 8) * F(1.0 / 16777216.0)
    with oracle.new_scene() as s:
        s.begin_medium(float(density), s.texture_solid(.2, .4, .9))
        s.sphere(tuple(map(float, c)), float(r), s.dielectric(1.5))
        s.end_medium()
        s.build()
        o, d = _rays(5000, 31)
        h = _trace(s, o, d)
    neg_inv = F(-1.0) / density
    hit_distance = neg_inv * F(np.log10(np.float64(xi)))
    n_hit = n_miss = 0
    for i in range(len(o)):
        rec1 = _sphere_hit(o[i], d[i], c, r, F(-np.inf), F(np.inf))
        rec2 = _sphere_hit(o[i], d[i], c, r, rec1[0] + F(0.0001), F(np.inf)) if rec1 else None
        want = None
        if rec1 and rec2:
            t1, t2 = max(rec1[0], F(0.001)), min(rec2[0], F(np.inf))
            if not t1 >= t2:
                t1 = max(t1, F(0.0))
                length = np.sqrt(_dot(d[i], d[i]))
                inside = (t2 - t1) * length
                margin = abs(float(hit_distance) - float(inside))
                if margin < 1e-4 * float(inside):
                    continue                                   # at the threshold: libm's log10 may decide
                if not hit_distance > inside:
                    want = t1 + hit_distance / length
        if want is None:
            assert h["prim_id"][i] < 0, i
            n_miss += 1
        else:
            assert h["prim_id"][i] == 0
            np.testing.assert_allclose(h["t"][i], want, rtol=1e-6)
            np.testing.assert_allclose(h["p"][i], o[i] + h["t"][i] * d[i], rtol=1e-6, atol=1e-6)
            np.testing.assert_array_equal(h["normal"][i], _v(1, 0, 0))
            assert h["front_face"][i] == 1
            n_hit += 1
    assert n_hit > 300 and n_miss > 100
