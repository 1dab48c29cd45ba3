"""GPU parity tests: the CUDA backend (through its C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): closest-hit primitive ids bit-exact except documented exact ties (none
occur: the backend implements the reference list's "last tested wins" rule exactly); hit t / p / normal
bit-exact (stronger than the 1e-5 asked); uv within 1e-5 (CUDA acosf/atan2f differ from glibc by <= 2 ulp);
images bit-exact where no libm call is on the path (Cornell box), RMSE-bounded otherwise.
"""
import numpy as np
import pytest

import raytracer_weekend_b200 as rtw
from conftest import bits

pytestmark = pytest.mark.gpu

SCENES = [("cornell-box", 1.0), ("jumpy-balls", 16 / 9), ("cow-lambert-metal", 16 / 9), ("wavefront-cow-obj", 16 / 9),
          ("simple-triangle", 16 / 9), ("two-spheres", 16 / 9), ("two-perlin-spheres", 16 / 9), ("earth", 16 / 9),
          ("simple-light", 16 / 9), ("monument-earth", 16 / 9), ("stress:3000:400", 16 / 9), ("smokey-cornell-box", 1.0),
          ("book2-final-scene", 1.0), ("animated-book2-final-scene", 1.0)]


def assert_hits_equal(hg, ho, what, medium_ids=()):
    if len(medium_ids):
        # ConstantMedium hits go through log10f (CUDA <= 2 ulp vs glibc): t / p to 1e-5, everything else exact
        med = np.isin(ho["prim_id"], medium_ids) | np.isin(hg["prim_id"], medium_ids)
        assert (hg["prim_id"][med] == ho["prim_id"][med]).mean() > 0.999
        both = med & (hg["prim_id"] == ho["prim_id"])
        np.testing.assert_allclose(hg["t"][both], ho["t"][both], rtol=1e-5, err_msg=what)
        np.testing.assert_allclose(hg["p"][both], ho["p"][both], rtol=1e-5, atol=1e-3, err_msg=what)
        assert np.array_equal(hg["normal"][both], ho["normal"][both]) and (hg["front_face"][both] == 1).all()
        hg, ho = hg[~med], ho[~med]
    same_id = hg["prim_id"] == ho["prim_id"]
    assert same_id.all(), f"{what}: {np.count_nonzero(~same_id)} closest-hit id mismatches of {len(hg)}"
    assert np.array_equal(bits(hg["t"]), bits(ho["t"])), f"{what}: t differs"
    assert np.array_equal(bits(hg["p"]), bits(ho["p"])), f"{what}: p differs"
    assert np.array_equal(bits(hg["normal"]), bits(ho["normal"])), f"{what}: normal differs"
    assert np.array_equal(hg["front_face"], ho["front_face"]), f"{what}: front_face differs"
    assert np.array_equal(hg["material_id"], ho["material_id"]), f"{what}: material differs"
    np.testing.assert_allclose(hg["u"], ho["u"], rtol=1e-5, atol=2e-6, err_msg=what)   # tolerance of the north star
    np.testing.assert_allclose(hg["v"], ho["v"], rtol=1e-5, atol=2e-6, err_msg=what)


def test_smoke_entry():
    import __graft_entry__ as g

    g.smoke()


@pytest.mark.parametrize("scene,aspect", SCENES)
def test_trace_parity_on_captured_ray_batches(gpu, oracle, scene, aspect):
    w, h = 192, int(round(192 / aspect))
    with rtw.Scene.from_name(gpu, scene, w / h, seed=1) as sg, rtw.Scene.from_name(oracle, scene, w / h, seed=1) as so:
        cam = sg.cameras[-1]
        media = [i for i in range(sg.num_prims) if sg.prim_info(i)[0] >= 6] if "cornell" in scene or "book2" in scene else []
        for bounce in (0, 1, 2, 4):
            rays = oracle.capture_rays(so, cam, w, h, 11, bounce, bounce)  # sample index = bounce: different jitter
            ho = so.trace_closest(rays)
            assert_hits_equal(sg.trace_closest(rays, rtw.RTW_TRACE_BVH), ho, f"{scene} bounce {bounce} LBVH", media)
            assert_hits_equal(sg.trace_closest(rays, rtw.RTW_TRACE_BRUTE), ho, f"{scene} bounce {bounce} brute force", media)


def test_trace_parity_on_adversarial_rays(gpu, oracle):
    """axis-aligned rays, rays starting on surfaces, along box edges and diagonals (exact-t ties), finite windows"""
    rs = np.random.RandomState(5)
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0) as sg, rtw.Scene.from_name(oracle, "cornell-box", 1.0) as so:
        o = rs.uniform(1, 554, (20000, 3)).astype(np.float32)
        d = rs.uniform(-1, 1, (20000, 3)).astype(np.float32)
        d[:4000, 0] = 0                                    # 1/d = inf on one axis
        d[4000:6000] = np.eye(3, dtype=np.float32)[rs.randint(0, 3, 2000)] * rs.choice([-1, 1], (2000, 1))
        o[6000:8000, 1] = 0.0                              # origins on the floor plane
        o[8000:9000] = (278, 278, -800)
        d[8000:9000] = np.array([[0, 0, 1]], np.float32) + rs.uniform(-.35, .35, (1000, 3)).astype(np.float32) * [1, 1, 0]
        o[9000:9500] = (0, 0, 0)                           # corner to corner
        d[9000:9500] = (555, 555, 555)
        rays = rtw.make_rays(o, d)
        rays["t_max"][10000:12000] = rs.uniform(1, 800, 2000)
        rays["t_min"][12000:13000] = rs.uniform(0, 300, 1000)
        keep = np.isfinite(so.trace_closest(rays)["t"])    # in-plane NaN hits are undefined behaviour of the reference
        ho = so.trace_closest(rays[keep])
        assert_hits_equal(sg.trace_closest(rays[keep]), ho, "cornell adversarial")
    with gpu.new_scene() as sg, oracle.new_scene() as so:   # stacked coincident / touching primitives
        for s in (sg, so):
            m = s.lambertian_rgb(.5, .5, .5)
            for _ in range(3):
                s.xy_rect(0, 1, 0, 1, 2, m)
            s.cuboid((0, 0, 3), (1, 1, 4), m)
            s.cuboid((1, 0, 3), (2, 1, 4), m)               # shares the x = 1 face
            s.sphere((0.5, 0.5, 6), 0.5, m)
            s.sphere((0.5, 0.5, 7), 0.5, m)                 # touches the first sphere at z = 6.5
            s.triangles([[0, 0, 8, 1, 0, 8, 0, 1, 8], [1, 0, 8, 1, 1, 8, 0, 1, 8]], m)  # shared diagonal edge
            s.build()
        o = (rs.uniform(-.5, 2.5, (30000, 3)) * [1, 1, 0] + [0, 0, -1]).astype(np.float32)
        d = np.tile(np.array([[0, 0, 1]], np.float32), (30000, 1))
        # exact-t ties without lying in any face plane (that is the reference's NaN quirk, see
        # test_oracle_kat.py::test_rectangles): slanted dyadic rays o = p - 4 d through lattice points p
        # of the planes z = 2 (three coincident rects) and z = 3 (the two cuboid bottoms incl. their shared
        # edge x = 1, where XY@z0 of both boxes and the coincident YZ faces all report t = 4)
        lat = np.round(rs.uniform(-.5, 2.5, (6000, 2)) * 4) / 4
        for k, zt, dd in ((slice(0, 3000), 2.0, (0.5, 0.25, 1.0)), (slice(3000, 6000), 3.0, (-0.5, 0.25, 1.0))):
            dd = np.array(dd, np.float32)
            p = np.concatenate([lat[k], np.full((3000, 1), zt)], axis=1).astype(np.float32)
            o[k] = p - 4 * dd
            d[k] = dd
        o[6000:7000] = (3, .5, 3.5)
        d[6000:7000] = (-1, 0, 0)
        rays = rtw.make_rays(o, d)
        ho = so.trace_closest(rays)
        assert np.isfinite(ho["t"][ho["prim_id"] >= 0]).all()
        hg = sg.trace_closest(rays)
        assert_hits_equal(hg, ho, "coincident primitives")
        on_rects = ho["t"][:3000] == 4.0
        assert on_rects.sum() > 200 and (ho["prim_id"][:3000][on_rects] == 2).all()   # the last coincident rect wins
        edge = (lat[3000:6000, 0] == 1.0) & (lat[3000:6000, 1] > 0) & (lat[3000:6000, 1] < 1)
        assert edge.sum() > 10 and (ho["prim_id"][3000:6000][edge] == 14).all()       # 4-way tie: highest id


def test_render_cornell_bit_exact(gpu, oracle):
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0) as sg, rtw.Scene.from_name(oracle, "cornell-box", 1.0) as so:
        cam = sg.cameras[0]
        for w, h, spp, slices, pool in ((96, 96, 12, 1, 0), (96, 96, 12, 4, 4096), (128, 64, 7, 3, 1 << 16), (33, 47, 5, 5, 96)):
            p = sg.params(w, h, spp, seed=5, slices=slices, pool_size=pool)
            ag, stg = sg.render(cam, p)
            ao, sto = so.render(cam, p)
            assert stg.segments == sto.segments and stg.paths == sto.paths == w * h * spp
            assert np.array_equal(bits(ag), bits(ao)), (w, h, spp, slices, pool)


def test_render_cornell_full_resolution_bit_exact(gpu, oracle):
    """BASELINE config C2's frame size (800x800) at 2 spp against the oracle, every pixel bit for bit."""
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0) as sg, rtw.Scene.from_name(oracle, "cornell-box", 1.0) as so:
        p = sg.params(800, 800, 2, seed=2024, slices=2)
        ag, stg = sg.render(sg.cameras[0], p)
        ao, sto = so.render(so.cameras[0], p)
        assert stg.segments == sto.segments
        assert np.array_equal(bits(ag), bits(ao))


@pytest.mark.parametrize("scene,aspect", [("jumpy-balls", 16 / 9), ("cow-lambert-metal", 16 / 9), ("monument-earth", 16 / 9),
                                          ("two-perlin-spheres", 16 / 9), ("simple-light", 16 / 9), ("stress:3000:400", 16 / 9),
                                          ("simple-triangle", 16 / 9), ("earth", 16 / 9),
                                          ("smokey-cornell-box", 1.0), ("book2-final-scene", 1.0)])
def test_render_statistical_parity(gpu, oracle, scene, aspect):
    """Scenes with sinf / acosf / atan2f on the path: same stream, same control flow except where a
    <= 2 ulp libm difference flips a checker cell / texel / Perlin value.  Stated bound: at equal spp the
    images agree to RMSE <= 2% of the mean radiance and >= 97% of the pixels agree to 1e-3 relative."""
    w, h, spp = 160, int(round(160 / aspect)), 8
    with rtw.Scene.from_name(gpu, scene, w / h, seed=1) as sg, rtw.Scene.from_name(oracle, scene, w / h, seed=1) as so:
        p = sg.params(w, h, spp, seed=77, slices=2)
        ag, stg = sg.render(sg.cameras[0], p)
        ao, sto = so.render(so.cameras[0], p)
        assert abs(int(stg.segments) - int(sto.segments)) <= 0.01 * sto.segments
        # a 1-ulp sinf difference changes a Perlin / checker attenuation by an ulp without changing the path:
        # count pixels that agree to 1e-3 relative (paths that really diverged differ by far more)
        same = np.isclose(ag, ao, rtol=1e-3, atol=1e-5).all(axis=2).mean()
        rmse = float(np.sqrt(np.mean((ag - ao) ** 2)))
        assert rmse <= 0.02 * float(np.mean(ao)) + 1e-6, (rmse, float(np.mean(ao)))
        assert same >= 0.97, same


def test_image_is_independent_of_pool_partition_and_counters(gpu):
    with rtw.Scene.from_name(gpu, "jumpy-balls", 16 / 9, seed=3) as s:
        cam = s.cameras[0]
        w, h, spp = 200, 112, 6
        ref, st0 = s.render(cam, s.params(w, h, spp, seed=9, slices=3))
        for pool in (64, 5000, 1 << 18):
            a, st = s.render(cam, s.params(w, h, spp, seed=9, slices=3, pool_size=pool))
            assert np.array_equal(bits(a), bits(ref)) and st.segments == st0.segments
        a, st = s.render(cam, s.params(w, h, spp, seed=9, slices=3, flags=rtw.RTW_RENDER_COUNT_TRAVERSAL | rtw.RTW_RENDER_TIME_KERNELS))
        assert np.array_equal(bits(a), bits(ref))
        assert st.node_visits > st.segments and st.prim_tests > 0 and st.prim_bytes > 0 and st.ms_traverse > 0
        for parts, ts in ((2, 32), (3, 16), (8, 32), (5, 24)):   # 24: the non-power-of-two tile path
            total = np.zeros_like(ref)
            seg = 0
            for r in range(parts):
                a, st = s.render(cam, s.params(w, h, spp, seed=9, slices=3, part_rank=r, part_count=parts, tile_size=ts))
                assert np.count_nonzero(total[a != 0]) == 0       # disjoint pixel sets
                total += a
                seg += st.segments
            assert np.array_equal(bits(total), bits(ref)) and seg == st0.segments
        a, _ = s.render(cam, s.params(w, h, spp, seed=9, slices=3, sample_begin=0, sample_end=2))
        b, _ = s.render(cam, s.params(w, h, spp, seed=9, slices=3, sample_begin=2, sample_end=6))
        np.testing.assert_allclose(a + b, ref, rtol=2e-6, atol=1e-7)   # sample slices: same samples, other summation tree
        z, st = s.render(cam, s.params(w, h, spp, seed=9, sample_begin=3, sample_end=3))
        assert st.segments == 0 and not z.any()


def test_bvh_structure(gpu):
    """every primitive slot is reachable exactly once; every child box lies inside its parent's; leaves are
    contiguous slot ranges; the instance / type order inside a leaf is sorted"""
    for scene in ("cornell-box", "cow-lambert-metal", "jumpy-balls", "stress:2000:300"):
        with rtw.Scene.from_name(gpu, scene, 16 / 9, seed=1) as s:
            nodes, slot_ids, root = s.get_bvh()
            n = s.num_prims
            assert sorted(slot_ids.tolist()) == list(range(n))
            seen = np.zeros(n, np.int32)
            stack = [(0, root[:3], root[3:])]
            visited = 0
            while stack:
                pair, lo, hi = stack.pop()
                visited += 1
                for rec in nodes[2 * pair:2 * pair + 2]:
                    if rec["meta"] == 0 and rec["link"] < 0 and np.isinf(rec["bmin"]).any():
                        continue                                       # the empty sibling of a flat scene
                    assert (rec["bmin"] >= lo - 1e-6 * np.abs(lo)).all() and (rec["bmax"] <= hi + 1e-6 * np.abs(hi)).all()
                    if rec["link"] >= 0:
                        stack.append((int(rec["link"]), rec["bmin"], rec["bmax"]))
                    else:
                        first, cnt = ~int(rec["link"]), int(rec["meta"])
                        assert 1 <= cnt <= 32 and first + cnt <= n
                        seen[first:first + cnt] += 1
                        metas = [s.prim_info(int(i))[1] * 8 + s.prim_info(int(i))[0] for i in slot_ids[first:first + cnt]]
                        assert metas == sorted(metas)
            assert (seen == 1).all(), scene
            assert s.build_stats.num_prims == n and s.build_stats.max_depth < 90


def _internal_area(scene):
    """sum of SA over the internal nodes of the built tree / SA(root): the expected pair visits of a random ray"""
    nodes, _, root = scene.get_bvh()
    def area(lo, hi):
        d = np.maximum(hi.astype(np.float64) - lo.astype(np.float64), 0)
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]
    total, depth_max, stack = area(root[:3], root[3:]), 0, [(0, 1)]
    while stack:
        pair, depth = stack.pop()
        depth_max = max(depth_max, depth)
        for rec in nodes[2 * pair:2 * pair + 2]:
            if rec["link"] >= 0:
                total += area(rec["bmin"], rec["bmax"])
                stack.append((int(rec["link"]), depth + 1))
    return total / area(root[:3], root[3:]), depth_max


def test_tree_rotations_lower_the_sah_cost_and_change_no_hit(gpu, oracle, monkeypatch):
    """rtw_bvh.cu: k_rotate.  The rotated tree has the same leaves and boxes that still contain them (test_bvh_structure runs
    on it), a clearly lower expected number of pair visits than the plain Karras tree (RTW_ROTATE=0), a depth the
    traversal stack holds — and the same closest hits and the same frame, bit for bit: world.hit does not depend on the
    shape of the tree."""
    for scene, gain in (("cow-lambert-metal", 0.80), ("jumpy-balls", 0.60), ("stress:3000:400", 0.99)):
        res = {}
        for passes in ("0", "4"):
            monkeypatch.setenv("RTW_ROTATE", passes)   # read by rtw_build
            with rtw.Scene.from_name(gpu, scene, 16 / 9, seed=3) as sg, rtw.Scene.from_name(oracle, scene, 16 / 9, seed=3) as so:
                cost, depth = _internal_area(sg)
                assert depth + 2 <= 96 and sg.build_stats.max_depth + 2 <= 96
                rays = oracle.capture_rays(so, so.cameras[0], 160, 90, 11, 0, 1)
                hits = sg.trace_closest(rays)
                a, st = sg.render(sg.cameras[0], sg.params(96, 54, 4, seed=5, slices=2))
                res[passes] = (cost, hits, bits(a).copy(), st.segments)
        monkeypatch.delenv("RTW_ROTATE")
        assert res["4"][0] <= gain * res["0"][0], (scene, res["0"][0], res["4"][0])
        assert_hits_equal(res["4"][1], res["0"][1], f"rotations {scene}")
        assert np.array_equal(res["4"][2], res["0"][2]) and res["4"][3] == res["0"][3]


def test_block_cache_recycles_and_trims(gpu):
    """rtw_mem.cu: the blocks of a destroyed scene are handed to the next one (dirty: nothing may rely on zeroed memory) and
    rtw_trim_memory gives them back to the driver; frames are the same before and after."""
    frames = []
    for rep in range(3):
        with rtw.Scene.from_name(gpu, "jumpy-balls", 16 / 9, seed=3) as s:
            a, st = s.render(s.cameras[0], s.params(80, 45, 4, seed=2, slices=2))
            frames.append(bits(a).copy())
        if rep == 1:
            assert gpu.trim_memory() >= 1   # the wavefront pool alone is > 1 MiB
            assert gpu.trim_memory() == 0   # nothing left to release
    assert np.array_equal(frames[0], frames[1]) and np.array_equal(frames[0], frames[2])


def test_resolve_rgb8_matches_reference_tonemap(gpu, oracle):
    rs = np.random.RandomState(0)
    acc = (rs.uniform(0, 3, (64, 48, 3)) ** 3).astype(np.float32) * 16
    acc[0, 0] = (np.nan, -1.0, 1e30)
    with gpu.new_scene() as sg, oracle.new_scene() as so:
        assert np.array_equal(sg.resolve_rgb8(acc, 16), so.resolve_rgb8(acc, 16))


def test_state_errors(gpu):
    with gpu.new_scene() as s:
        m = s.lambertian_rgb(.5, .5, .5)
        s.sphere((0, 0, -2), 1.0, m)
        with pytest.raises(rtw.RtwError, match="not built"):
            s.render(rtw.camera_new((0, 0, 0), (0, 0, -1), (0, 1, 0), 40, 1.0), s.params(8, 8, 1))
        s.build()
        with pytest.raises(rtw.RtwError, match="already built"):
            s.sphere((0, 0, -2), 1.0, m)
        cam = rtw.camera_new((0, 0, 0), (0, 0, -1), (0, 1, 0), 40, 1.0)
        with pytest.raises(rtw.RtwError, match=">= 2"):
            s.render(cam, s.params(1, 8, 1))
        with pytest.raises(rtw.RtwError, match="part_rank"):
            s.render(cam, s.params(8, 8, 1, part_rank=3, part_count=2))
        a, st = s.render(cam, s.params(8, 8, 2, background=(0.5, 0.25, 1.0)))   # single primitive scene
        assert st.paths == 128 and a.shape == (8, 8, 3) and np.isfinite(a).all() and a.max() > 0
        assert s.trace_closest(np.zeros(0, rtw.RAY_DTYPE)).shape == (0,)        # empty batch


def test_large_scene_mismatches_are_only_ill_conditioned_reference_hits(gpu, oracle):
    """SURVEY.md §8a exception 2 (cull disagreements), measured.  Seen from 260 units away the reference's f32
    sphere formula (spherical.rs:27-31: half_b^2 - a*c with both terms ~1e5) is rounding noise for a sphere
    of radius 0.05-0.25: its flat list reports "hits" whose point is 1.2-2 radii from the centre.  Any BVH —
    the reference's own BvhNode included — culls those because the ray misses the sphere's box.  Every
    difference between the LBVH and the flat list must be of that kind; brute force must equal the oracle."""
    name = "stress:60000:3000"
    with rtw.Scene.from_name(gpu, name, 16 / 9, seed=2024) as sg, rtw.Scene.from_name(oracle, name, 16 / 9, seed=2024) as so:
        rs = np.random.RandomState(0)
        n = 6000
        o = np.tile([[0, 0, -260]], (n, 1)).astype(np.float32)
        d = (np.array([[0, 0, 1]]) + rs.uniform(-.3, .3, (n, 3)) * [1, 1, 0]).astype(np.float32)
        rays = rtw.make_rays(o, d)
        ho = so.trace_closest(rays)
        assert_hits_equal(sg.trace_closest(rays, rtw.RTW_TRACE_BRUTE), ho, "brute force vs flat list")
        hg = sg.trace_closest(rays, rtw.RTW_TRACE_BVH)
        diff = (hg["prim_id"] != ho["prim_id"]) | (bits(hg["t"]) != bits(ho["t"]))
        nlen = np.linalg.norm(ho["normal"], axis=1)
        is_sphere = (ho["prim_id"] >= 0) & (ho["prim_id"] < 60000)
        bogus = is_sphere & (np.abs(nlen - 1.0) > 1e-3)           # hit point not on the sphere: |(p-c)/r| != 1
        assert not (diff & ~bogus).any(), "a well-conditioned hit was lost"
        assert diff.mean() < 0.02
        # where the oracle's hit is sound, everything is bit-exact
        ok = ~bogus & ~diff
        assert np.array_equal(bits(hg["p"][ok]), bits(ho["p"][ok])) and np.array_equal(bits(hg["normal"][ok]), bits(ho["normal"][ok]))
        print(f"cull disagreements: {int(diff.sum())} of {n} rays ({int(bogus.sum())} ill-conditioned oracle hits)")


def test_render_frames_keeps_the_scene_resident_and_matches_frame_by_frame(gpu, oracle):
    """rtw_render_frames (scenes.rs:622-667 + main.rs:48-95): frame i = cameras[i], seed + i; callbacks in frame order;
    bit-identical to one rtw_render per frame and to the oracle's frames (Cornell: no libm on the device path)."""
    import threading
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0, seed=1) as sg, rtw.Scene.from_name(oracle, "cornell-box", 1.0, seed=1) as so:
        cams = [sg.cameras[0], rtw.camera_new((278, 278, -700), (278, 278, 0), (0, 1, 0), 40.0, 1.0),
                rtw.camera_new((100, 300, -600), (278, 278, 0), (0, 1, 0), 45.0, 1.0), sg.cameras[0]]
        p = sg.params(48, 40, 5, seed=21, slices=2)
        got, order, threads = {}, [], set()

        def on_frame(i, accum, st):
            order.append(i)
            threads.add(threading.get_ident())
            got[i] = (accum, st["segments"])

        assert sg.render_frames(cams, p, on_frame) == 4
        assert order == [0, 1, 2, 3]
        assert threading.get_ident() not in threads          # delivered on the helper thread, overlapping the next render
        for i, c in enumerate(cams):
            a, st = sg.render(c, sg.params(48, 40, 5, seed=21 + i, slices=2))
            assert np.array_equal(bits(a), bits(got[i][0])) and st.segments == got[i][1]
            ao, sto = so.render(c, so.params(48, 40, 5, seed=21 + i, slices=2))
            assert np.array_equal(bits(ao), bits(got[i][0])) and sto.segments == got[i][1]
        assert not np.array_equal(got[0][0], got[3][0])      # same camera, another seed
        # stopping: the frames already in flight (at most two) are still delivered, later ones are not rendered
        seen = []
        n = sg.render_frames(cams * 3, p, lambda i, a, st: (seen.append(i), False)[1])
        assert 1 <= n <= 3 and seen == list(range(n))
        assert sg.render_frames(cams, p, None) == 4          # no callback: frames rendered and dropped
        with pytest.raises(rtw.RtwError):
            sg.render_frames(cams, sg.params(1, 1, 1), on_frame)


def test_tail_of_the_frame_switches_to_queues_and_stays_bit_exact(gpu, oracle):
    """Identity slot mapping while work items remain, compaction queues afterwards (rtw_render.cu): frames far smaller
    than, equal to and larger than the pool, all against the oracle."""
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0, seed=1) as sg, rtw.Scene.from_name(oracle, "cornell-box", 1.0, seed=1) as so:
        cam = sg.cameras[0]
        for (w, h, spp, pool, slices) in [(8, 8, 1, 0, 0), (33, 17, 3, 64, 2), (64, 64, 4, 4096, 1), (40, 40, 7, 1 << 20, 3)]:
            pg = sg.params(w, h, spp, seed=9, pool_size=pool, slices=slices)
            ag, stg = sg.render(cam, pg)
            ao, sto = so.render(cam, so.params(w, h, spp, seed=9, slices=stg.slices))
            assert stg.segments == sto.segments and stg.paths == w * h * spp
            assert np.array_equal(bits(ag), bits(ao)), (w, h, spp, pool)


@pytest.mark.parametrize("env,value,record_bytes", [("RTW_COMPACT", "1", 32.0), ("RTW_WIDE", "1", 128.0), ("RTW_FLAT", "0", 64.0)])
def test_alternative_node_records_keep_parity(gpu, oracle, monkeypatch, env, value, record_bytes):
    """Compact 32-byte pairs (16-bit boxes on the scene grid; what rtw_build picks for >= 2^20 primitives) and the
    experimental 4-wide walk, forced on for small scenes: closest hits and images stay what the oracle says — the
    quantised boxes contain the exact ones, so culling stays conservative.  RTW_FLAT=0: one-leaf scenes through the
    general kernel instead of the flat-scene kernel they normally take."""
    monkeypatch.setenv(env, value)   # read by rtw_build (compact) / rtw_render, rtw_trace_closest (wide, flat)
    for scene, aspect in (("cow-lambert-metal", 16 / 9), ("stress:3000:400", 16 / 9), ("cornell-box", 1.0)):
        with rtw.Scene.from_name(gpu, scene, aspect, seed=3) as sg, rtw.Scene.from_name(oracle, scene, aspect, seed=3) as so:
            cam = sg.cameras[0]
            for bounce in (0, 1):
                rays = oracle.capture_rays(so, cam, 160, 90, 11, 0, bounce)
                hg, ho = sg.trace_closest(rays), so.trace_closest(rays)
                if scene.startswith("stress"):   # ill-conditioned reference hits outside the sphere's box
                    assert (hg["prim_id"] != ho["prim_id"]).mean() < 0.01
                else:
                    assert_hits_equal(hg, ho, f"{env} {scene} bounce {bounce}")
            p = sg.params(64, 36, 3, seed=5, slices=1, flags=rtw.RTW_RENDER_COUNT_TRAVERSAL)
            ag, stg = sg.render(cam, p)
            if env == "RTW_COMPACT" and scene != "cornell-box":
                # (the flat Cornell scene has an empty right child, which the compact encoding refuses: it keeps its
                # fp32 pair; the counting variant of the 4-wide walk does not exist: it counts the pair walk)
                assert stg.node_record_bytes == record_bytes
            if scene == "cornell-box":
                ao, sto = so.render(cam, so.params(64, 36, 3, seed=5, slices=1))
                assert np.array_equal(bits(ag), bits(ao)) and stg.segments == sto.segments


def test_trace_parity_on_a_million_ray_batches(gpu, oracle):
    """1 M camera rays + 1 M second-bounce rays captured from the oracle, closest hits through
    rtw_trace_closest against the oracle's canonical-order flat list: ids / t / p / normal bit for bit, uv 1e-5."""
    with rtw.Scene.from_name(gpu, "cow-lambert-metal", 16 / 9, seed=3) as sg, \
            rtw.Scene.from_name(oracle, "cow-lambert-metal", 16 / 9, seed=3) as so:
        cam = sg.cameras[0]
        for bounce in (0, 1):
            rays = oracle.capture_rays(so, cam, 1366, 768, 2024, 0, bounce)          # 1 049 088 rays
            assert len(rays) > 1_000_000
            hg, ho = sg.trace_closest(rays), so.trace_closest(rays)
            assert (ho["prim_id"] >= 0).sum() > (100_000 if bounce == 0 else 20_000)   # ended paths yield null rays
            assert_hits_equal(hg, ho, f"cow, 1M rays, bounce {bounce}")
            brute = sg.trace_closest(rays[:50_000], mode=rtw.RTW_TRACE_BRUTE)          # the GPU's own flat list
            assert np.array_equal(brute["prim_id"], hg["prim_id"][:50_000]) and np.array_equal(bits(brute["t"]), bits(hg["t"][:50_000]))


def test_render_against_committed_golden_frames(gpu):
    """The CUDA path against tests/golden/oracle_renders_v1.npz (frames the oracle rendered once, committed): bit for
    bit where no libm call is on the device path (Cornell box), within the stated statistical bound elsewhere
    (sinf / acosf / atan2f / log10f differ from glibc by <= 2 ulp, which can flip a checker cell / texel / medium
    scatter for a rare path: RMSE <= 5 % of the mean radiance and >= 93 % of the pixels equal to 1e-3 relative — the
    frames are tiny, one diverged path weighs 1 / 4000).  No oracle code runs in this test."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_renders_v1.npz"))
    for scene in sorted({k.split("/")[0] for k in g.files}):
        w, h, spp, slices, segments = (int(x) for x in g[f"{scene}/meta"])
        want = g[f"{scene}/accum"]
        with rtw.Scene.from_name(gpu, scene, w / h, seed=1) as s:
            got, st = s.render(s.cameras[0], s.params(w, h, spp, seed=4242, slices=slices))
        if scene == "cornell-box":
            assert st.segments == segments and np.array_equal(bits(got), bits(want))
            continue
        assert abs(int(st.segments) - segments) <= 0.02 * segments, scene
        rmse = float(np.sqrt(np.mean((got - want) ** 2)))
        same = np.isclose(got, want, rtol=1e-3, atol=1e-5).all(axis=2).mean()
        print(f"{scene}: rmse / mean = {rmse / float(np.mean(want)):.5f}, identical pixels = {same:.4f}")
        assert rmse <= 0.05 * float(np.mean(want)) + 1e-6, (scene, rmse)
        assert same >= 0.93, (scene, same)


def test_console_app_backend_cuda_end_to_end(gpu, tmp_path):
    """console_app (main.rs:28-95 mirror) with --backend cuda: PNG per camera through rtw_render_frames, identical to
    a direct render + the reference tonemap; the --progress-out stream decodes to the same frame (postcard + COBS)."""
    import os
    import struct
    import subprocess
    from PIL import Image
    exe = os.path.join(rtw.PKG_DIR, "bin", "console_app")
    out = tmp_path / "render"
    prog = tmp_path / "progress.bin"
    r = subprocess.run([exe, "-w", "48", "-a", "1.0", "-s", "5", "--seed", "9", "--backend", "cuda", "--out-dir", str(out),
                        "--progress-out", str(prog), "cornell-box"], stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    assert "1 frame(s)" in r.stderr
    img = np.asarray(Image.open(out / "image_0000.png").convert("RGB"))
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0, seed=9) as s:
        accum, _ = s.render(s.cameras[0], s.params(48, 48, 5, seed=9))
        want = s.resolve_rgb8(accum, 5)
    assert img.shape == (48, 48, 3) and np.array_equal(img, want)
    frames = prog.read_bytes().split(b"\x00")
    assert frames[-1] == b"" and len(frames) == 1 + 48 * 48 + 1 + 1

    def cobs_decode(fr):
        o, i = bytearray(), 0
        while i < len(fr):
            c = fr[i]
            o += fr[i + 1:i + c]
            i += c
            if c != 0xFF and i < len(fr):
                o.append(0)
        return bytes(o)

    assert cobs_decode(frames[0]) == bytes([0]) + struct.pack("<III", 48, 48, 5)
    tag, row, col, red, green, blue = struct.unpack("<BIIfff", cobs_decode(frames[1]))
    assert (tag, row, col) == (1, 47, 0) and np.array_equal(np.float32([red, green, blue]), accum[0, 0])
    assert cobs_decode(frames[-2]) == bytes([2])


# ---- round 2: the parity gaps VERDICT r01 listed -----------------------------------------------------------------------
def _metal_glass_scene(s, n_small):
    """Solid-colour Lambertian + Metal + Dielectric spheres (RTiOW 'three spheres' plus a field of small ones): no
    texture lookup, no sinf / acosf / atan2f on the device path (uv is never needed), so the image must be bit-exact.
    n_small <= 26 keeps the scene one leaf (fused kernel); more goes through the LBVH + wavefront."""
    ground = s.lambertian_rgb(.5, .5, .5)
    s.sphere((0, -1000, 0), 1000.0, ground)
    s.sphere((0, 1, 0), 1.0, s.dielectric(1.5))
    s.sphere((0, 1, 0), -0.9, s.dielectric(1.5))                 # hollow glass: negative radius (scenes.rs:90-94)
    s.sphere((-4, 1, 0), 1.0, s.lambertian_rgb(.4, .2, .1))
    s.sphere((4, 1, 0), 1.0, s.metal(.7, .6, .5, 0.0))
    s.xz_rect(-3, 3, -3, 3, 6.0, s.diffuse_light_rgb(4, 4, 4))
    rs = np.random.RandomState(n_small)
    for i in range(n_small):
        c = (float(rs.uniform(-6, 6)), 0.2, float(rs.uniform(-5, 5)))
        k = i % 4
        m = (s.lambertian_rgb(*rs.uniform(0, 1, 3)) if k == 0 else s.metal(*rs.uniform(.5, 1, 3), float(rs.uniform(0, .5)))
             if k in (1, 2) else s.dielectric(1.5))
        s.sphere(c, 0.2, m)
    s.build()


@pytest.mark.parametrize("n_small", [8, 90])
def test_render_metal_and_dielectric_bit_exact(gpu, oracle, n_small):
    """material.rs:63-147: Metal (reflect + fuzz * random_in_unit_sphere) and Dielectric (Schlick, refract, the
    short-circuited draw) compared bit for bit — r01 only had them behind checker textures, i.e. statistically."""
    with gpu.new_scene() as sg, oracle.new_scene() as so:
        _metal_glass_scene(sg, n_small)
        _metal_glass_scene(so, n_small)
        cam = rtw.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 25.0, 1.5, aperture=0.1, focus_dist=10.0)
        rays = oracle.capture_rays(so, cam, 96, 64, 3, 0, 2)
        assert_hits_equal(sg.trace_closest(rays), so.trace_closest(rays), "metal-glass bounce 2")
        p = sg.params(96, 64, 16, seed=31, slices=2, background=(.7, .8, 1.0))
        ag, stg = sg.render(cam, p)
        ao, sto = so.render(cam, p)
        assert stg.fused == (1 if n_small == 8 else 0)
        assert stg.segments == sto.segments and stg.paths == sto.paths == 96 * 64 * 16
        assert np.array_equal(bits(ag), bits(ao))
        mats = {sg.prim_info(i)[2] for i in range(sg.num_prims)}
        assert len(mats) >= 5


def test_render_uvdebug_triangles_bit_exact(gpu, oracle):
    """texture.rs:97-104 (UVDebug) on triangles with default and with per-vertex uvs / normals (triangular.rs:42-73,
    315-323): barycentric arithmetic only, so the shaded image is bit-exact; scenes.rs:690 uses it."""
    def build(s):
        uvd = s.lambertian(s.texture_uvdebug())
        s.sphere((0, -1000, 0), 1000.0, s.lambertian_rgb(.5, .5, .5))
        s.triangles([[-5, 0, 5, 0, 7, 0, 5, 0, -5]], uvd)                                   # flat shaded, default uvs
        s.push_translation((0, 0, -6))
        s.triangles([[-4, 0, 0, 4, 0, 0, 0, 5, 0], [4, 0, 0, 4, 5, 0, 0, 5, 0]], uvd,
                    normals=[[0, 0, 1, 0, 0, 1, .3, 0, 1], [0, 0, 1, .2, .1, 1, .3, 0, 1]],
                    uvs=[[0, 0, 1, 0, .5, 1], [1, 0, 1, 1, .5, 1]])
        s.pop_transform()
        s.xz_rect(-4, 4, -4, 4, 12.0, s.diffuse_light(s.texture_uvdebug()))                 # emits (u, v, 0)
        s.build()
    with gpu.new_scene() as sg, oracle.new_scene() as so:
        build(sg)
        build(so)
        cam = rtw.camera_new((13, 3, 9), (0, 2.5, -2), (0, 1, 0), 40.0, 1.5)
        p = sg.params(90, 60, 12, seed=8, slices=3, background=(.1, .1, .12))
        ag, stg = sg.render(cam, p)
        ao, sto = so.render(cam, p)
        assert stg.segments == sto.segments and np.array_equal(bits(ag), bits(ao))
        assert ag[..., 0].max() > 0 and ag[..., 2].min() >= 0


@pytest.mark.parametrize("scene,w,h", [("cornell-box", 64, 64), ("cow-lambert-metal", 64, 36)])
def test_converged_image_with_independent_streams(gpu, oracle, scene, w, h):
    """north_star correctness clause 3: converged images match within a stated RMSE at equal spp, with INDEPENDENT
    random streams (every other render test shares the Philox stream with the oracle).  GPU(seed A) vs oracle(seed B)
    at 256 spp must be as close as two oracle renders with seeds B and C are to each other: RMSE(gpu_A, orc_B) <=
    1.25 x RMSE(orc_B, orc_C) (oracle-only triples give 0.99-1.03), measured on the tonemapped frame (main.rs:73-86: sqrt(mean), clamped — so that one
    firefly does not decide the test) and the mean radiance of the frames within 3 %."""
    spp = 256
    with rtw.Scene.from_name(gpu, scene, w / h, seed=1) as sg, rtw.Scene.from_name(oracle, scene, w / h, seed=1) as so:
        tone = lambda a: np.sqrt(np.clip(a / spp, 0.0, 1.0))  # noqa: E731
        ag, _ = sg.render(sg.cameras[0], sg.params(w, h, spp, seed=1001))
        # mode 2 = the reference's own structure (flat list + BvhNode, bvh.rs): same estimator, far fewer tests per ray
        ob, _ = oracle.render_ex(so, so.cameras[0], so.params(w, h, spp, seed=2002), mode=2)
        oc, _ = oracle.render_ex(so, so.cameras[0], so.params(w, h, spp, seed=3003), mode=2)
        rmse = lambda a, b: float(np.sqrt(np.mean((tone(a) - tone(b)) ** 2)))  # noqa: E731
        floor = rmse(ob, oc)
        got = max(rmse(ag, ob), rmse(ag, oc))
        print(f"{scene}: RMSE gpu-vs-oracle {got:.5f}, oracle-vs-oracle noise floor {floor:.5f}")
        assert floor > 0 and got <= 1.25 * floor, (got, floor)
        assert abs(float(tone(ag).mean()) - float(tone(ob).mean())) <= 0.03 * float(tone(ob).mean())
        assert not np.array_equal(ag, ob)


def test_million_primitive_scene_uses_compact_pairs_and_matches_brute_force(gpu):
    """C5's code path inside the suite: >= 2^20 primitives -> rtw_build picks the 32-byte compact pairs on its own;
    the LBVH walk must equal the GPU's own canonical-order flat list except where the reference's f32 sphere test is
    ill-conditioned (hit point not on the sphere: DESIGN.md, SURVEY.md §8a exception 2)."""
    with rtw.Scene.from_name(gpu, "stress:200000:100000", 16 / 9, seed=2024) as s:
        assert s.num_prims == 1_200_000
        rs = np.random.RandomState(1)
        n = 3000
        o = np.tile([[0, 0, -260]], (n, 1)).astype(np.float32)
        d = (np.array([[0, 0, 1]]) + rs.uniform(-.3, .3, (n, 3)) * [1, 1, 0]).astype(np.float32)
        cam_rays = rtw.make_rays(o, d)
        first = s.trace_closest(cam_rays, rtw.RTW_TRACE_BVH)
        hit = first["prim_id"] >= 0
        assert hit.sum() > 500
        # second generation: from the hit points into random directions (incoherent, deep in the scene)
        o2 = (first["p"][hit] + 1e-3 * first["normal"][hit]).astype(np.float32)
        d2 = rs.normal(size=o2.shape).astype(np.float32)
        for what, rays in (("camera", cam_rays), ("bounce", rtw.make_rays(o2, d2))):
            hb = s.trace_closest(rays, rtw.RTW_TRACE_BRUTE)
            hg = s.trace_closest(rays, rtw.RTW_TRACE_BVH)
            diff = (hg["prim_id"] != hb["prim_id"]) | (bits(hg["t"]) != bits(hb["t"]))
            nlen = np.linalg.norm(hb["normal"], axis=1)
            bogus = (hb["prim_id"] >= 0) & (hb["prim_id"] < 200000) & (np.abs(nlen - 1.0) > 1e-3)
            assert not (diff & ~bogus).any(), f"{what}: a well-conditioned hit was lost"
            assert diff.mean() < 0.02
            ok = ~diff
            assert np.array_equal(bits(hg["p"][ok]), bits(hb["p"][ok])) and np.array_equal(bits(hg["normal"][ok]), bits(hb["normal"][ok]))
        st = s.render_device_stats(s.cameras[0], s.params(160, 90, 2, seed=3, flags=rtw.RTW_RENDER_COUNT_TRAVERSAL))
        assert st.node_record_bytes == 32.0 and st.node_visits > st.segments > 160 * 90 * 2


@pytest.mark.parametrize("env", ["RTW_RAYSORT", "RTW_POOLED", "RTW_SHADE_SPLIT"])
def test_ray_reordering_and_pooled_leaf_tests_do_not_change_the_frame(gpu, oracle, monkeypatch, env):
    """world.hit does not depend on the order in which rays are traced (rtw_raysort.cuh: the rays of an iteration sorted
    by scene cell — on by default for hierarchies beyond the caches), nor on WHO tests a leaf (rtw_traverse.cuh:
    traverse_pooled, experiment), nor on which kernel restarts an ended path (RTW_SHADE_SPLIT: shade with bulk-copy staged
    state + k_wave_regen, experiment): forced on for small scenes, the frame keeps the bits of the default path (which the other
    tests pin to the oracle), including through the queue-mode tail of the frame (pool smaller than the frame) and with
    the traversal counters on."""
    for scene, aspect in (("stress:3000:400", 16 / 9), ("cow-lambert-metal", 16 / 9), ("jumpy-balls", 16 / 9)):
        with rtw.Scene.from_name(gpu, scene, aspect, seed=3) as sg:
            cam = sg.cameras[0]
            frames = {}
            for value in ("0", "1"):
                monkeypatch.setenv(env, value)
                for pool, flags in ((0, 0), (2048, 0), (4096, rtw.RTW_RENDER_COUNT_TRAVERSAL)):
                    a, st = sg.render(cam, sg.params(96, 54, 6, seed=5, slices=2, pool_size=pool, flags=flags))
                    if env == "RTW_RAYSORT":
                        assert st.ray_sort == int(value)
                    frames[(value, pool, flags)] = (bits(a).copy(), st.segments)
            ref = frames[("0", 0, 0)]
            for k, v in frames.items():
                assert np.array_equal(v[0], ref[0]) and v[1] == ref[1], (scene, env, k)


def test_million_primitive_scene_sorts_its_rays_by_default(gpu):
    """>= 2^20 primitives: compact pairs AND ray reordering are picked by the library itself; RTW_RAYSORT=0 must give the
    same frame (checked at a size the suite can afford)."""
    import os
    with rtw.Scene.from_name(gpu, "stress:200000:100000", 16 / 9, seed=2024) as s:
        p = s.params(160, 90, 3, seed=3, slices=1)
        a1, st1 = s.render(s.cameras[0], p)
        assert st1.ray_sort == 1 and st1.node_record_bytes == 32.0
        os.environ["RTW_RAYSORT"] = "0"
        try:
            a0, st0 = s.render(s.cameras[0], p)
        finally:
            del os.environ["RTW_RAYSORT"]
        assert st0.ray_sort == 0 and st0.segments == st1.segments
        assert np.array_equal(bits(a0), bits(a1))


def test_default_slices_do_not_depend_on_pool_or_partition(gpu):
    """with slices = 0 (auto) the slice count — which fixes the order of the float additions — used to
    follow the pool size and the tile partition, so the bits of a frame changed with the GPU count.  It now follows
    (width, height, samples, scene class) only."""
    for scene, aspect in (("cornell-box", 1.0), ("jumpy-balls", 16 / 9)):
        with rtw.Scene.from_name(gpu, scene, aspect, seed=3) as s:
            cam = s.cameras[0]
            w, h, spp = 128, int(round(128 / aspect)), 40
            ref, st0 = s.render(cam, s.params(w, h, spp, seed=9))
            assert st0.slices > 1
            for pool in (4096, 1 << 16):
                a, st = s.render(cam, s.params(w, h, spp, seed=9, pool_size=pool))
                assert st.slices == st0.slices and np.array_equal(bits(a), bits(ref))
            for parts in (2, 4, 8):
                total = np.zeros_like(ref)
                for r in range(parts):
                    a, st = s.render(cam, s.params(w, h, spp, seed=9, part_rank=r, part_count=parts))
                    assert st.slices == st0.slices
                    total += a
                assert np.array_equal(bits(total), bits(ref)), (scene, parts)


def test_fused_kernel_equals_the_wavefront(gpu, oracle, monkeypatch):
    """One-leaf scenes run in the fused persistent kernel (k_mega_flat); RTW_MEGA=0 sends them through the wavefront
    (k_wave_traverse_flat + k_wave_shade), RTW_FLAT=0 through the general LBVH walk: same bits, same segment count,
    all equal to the oracle."""
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0, seed=1) as sg, rtw.Scene.from_name(oracle, "cornell-box", 1.0, seed=1) as so:
        cam = sg.cameras[0]
        p = sg.params(80, 80, 9, seed=12, slices=3)
        a0, st0 = sg.render(cam, p)
        assert st0.fused == 1 and st0.launches <= 3
        ao, sto = so.render(cam, p)
        assert np.array_equal(bits(a0), bits(ao)) and st0.segments == sto.segments and st0.paths == 80 * 80 * 9
        for env in ("RTW_MEGA", "RTW_FLAT"):
            monkeypatch.setenv(env, "0")
            a1, st1 = sg.render(cam, p)
            monkeypatch.delenv(env)
            assert st1.fused == 0 and st1.iterations > 1
            assert np.array_equal(bits(a1), bits(a0)) and st1.segments == st0.segments
        # every slice / partition / sample-range combination of the fused kernel
        for kw in (dict(slices=1), dict(slices=9), dict(slices=4, part_rank=1, part_count=3, tile_size=16),
                   dict(slices=2, sample_begin=3, sample_end=7)):
            q = sg.params(80, 80, 9, seed=12, **kw)
            ag, stg = sg.render(cam, q)
            ao, sto = so.render(cam, q)
            assert np.array_equal(bits(ag), bits(ao)) and stg.segments == sto.segments, kw


def test_failed_render_leaks_nothing_and_the_scene_stays_usable(gpu, monkeypatch):
    """an error between the event / graph creates and their destroys leaked them and left the
    stream in capture mode.  Events, streams and the instantiated graph now live on the scene; a failure inside the
    graph capture or a failing launch returns an error, the next render works, and destroying the scene returns every
    handle (rtw_debug_live_handles counts them)."""
    base = gpu.live_handles()
    with rtw.Scene.from_name(gpu, "jumpy-balls", 16 / 9, seed=3) as s:      # hierarchy: wavefront + CUDA graph
        cam = s.cameras[0]
        p = s.params(64, 36, 3, seed=1, slices=1)
        monkeypatch.setenv("RTW_FAULT_INJECT", "capture")
        with pytest.raises(rtw.RtwError, match="graph capture"):
            s.render(cam, p)
        monkeypatch.delenv("RTW_FAULT_INJECT")
        held = gpu.live_handles()
        ref, st = s.render(cam, p)                                        # not stuck in capture mode
        assert st.paths == 64 * 36 * 3
        again, _ = s.render(cam, p)                                       # the cached graph is reused: no new handles
        assert np.array_equal(bits(again), bits(ref))
        assert gpu.live_handles() == held + 1                             # + the graph exec the failed call did not keep
        for _ in range(3):
            s.render(cam, s.params(64, 36, 3, seed=2, slices=1))
        assert gpu.live_handles() == held + 1
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0, seed=3) as s:       # one leaf: fused kernel
        monkeypatch.setenv("RTW_FAULT_INJECT", "launch")
        with pytest.raises(rtw.RtwError, match="CUDA error"):
            s.render(s.cameras[0], s.params(32, 32, 2, seed=1))
        monkeypatch.delenv("RTW_FAULT_INJECT")
        a, st = s.render(s.cameras[0], s.params(32, 32, 2, seed=1))
        assert st.paths == 32 * 32 * 2 and np.isfinite(a).all()
    assert gpu.live_handles() == base


def test_scene_clone_is_a_bit_identical_replica(gpu):
    """rtw_scene_clone: device buffers copied device to device, pointers rebased; the replica answers ray batches and
    renders exactly like the original (here onto the same device when the box has one GPU, else onto device 1)."""
    dev = 1 if gpu.device_count() > 1 else 0
    for scene, aspect in (("cow-lambert-metal", 16 / 9), ("cornell-box", 1.0), ("earth", 16 / 9)):
        with rtw.Scene.from_name(gpu, scene, aspect, seed=3) as s, s.clone(dev) as r:
            assert r.num_prims == s.num_prims and r.prim_info(0) == s.prim_info(0)
            n0, sl0, root0 = s.get_bvh()
            n1, sl1, root1 = r.get_bvh()
            assert np.array_equal(n0, n1) and np.array_equal(sl0, sl1) and np.array_equal(root0, root1)
            p = s.params(96, int(round(96 / aspect)), 4, seed=4, slices=2)
            a0, st0 = s.render(s.cameras[0], p)
            a1, st1 = r.render(r.cameras[0], p)
            assert np.array_equal(bits(a0), bits(a1)) and st0.segments == st1.segments


def _need_gpus(gpu, n):
    if gpu.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices (gpurun --gpus {n})")


@pytest.mark.parametrize("no_peer", ["0", "1"])
def test_render_over_two_gpus_through_the_c_abi(gpu, monkeypatch, no_peer):
    """rtw_render(..., gpus = N) — one process, a replica and a host thread per device, interleaved
    32x32 tiles, every device storing its pixels into the one frame over peer memory (no_peer = 1: staged peer copy +
    merge kernel).  The frame has the same bits as the single-GPU frame, default (auto) slices included."""
    _need_gpus(gpu, 2)
    monkeypatch.setenv("RTW_NO_PEER", no_peer)
    n = min(gpu.device_count(), 4)
    for scene, aspect in (("cornell-box", 1.0), ("cow-lambert-metal", 16 / 9)):
        with rtw.Scene.from_name(gpu, scene, aspect, seed=3) as s:
            cam = s.cameras[0]
            w, h = 200, int(round(200 / aspect))
            for kw in (dict(slices=0), dict(slices=1), dict(slices=5)):
                ref, st0 = s.render(cam, s.params(w, h, 6, seed=9, **kw))
                for g in sorted({2, n}):
                    a, st = s.render(cam, s.params(w, h, 6, seed=9, gpus=g, **kw))
                    assert st.gpus == g and st.segments == st0.segments and st.paths == st0.paths
                    assert np.array_equal(bits(a), bits(ref)), (scene, kw, g)
            frames = {}
            s.render_frames([cam, cam], s.params(w, h, 3, seed=5, gpus=2), lambda i, a, st: frames.__setitem__(i, (a, st["gpus"])))
            one, _ = s.render(cam, s.params(w, h, 3, seed=6))
            assert frames[1][1] == 2 and np.array_equal(bits(frames[1][0]), bits(one))


def test_gpus_argument_errors(gpu):
    with rtw.Scene.from_name(gpu, "cornell-box", 1.0) as s:
        cam = s.cameras[0]
        with pytest.raises(rtw.RtwError, match="device"):
            s.render(cam, s.params(16, 16, 1, gpus=gpu.device_count() + 1))
        if gpu.device_count() >= 2:
            with pytest.raises(rtw.RtwError, match="part_count"):
                s.render(cam, s.params(16, 16, 1, gpus=2, part_rank=0, part_count=2))


def test_console_app_gpus_flag(gpu, tmp_path):
    """console_app --backend cuda --gpus N (main.rs:15-26 + the new switch): the PNG is identical to --gpus 1."""
    import os
    import subprocess
    _need_gpus(gpu, 2)
    exe = os.path.join(rtw.PKG_DIR, "bin", "console_app")
    outs = []
    for g in (1, 2):
        out = tmp_path / f"g{g}"
        r = subprocess.run([exe, "-w", "120", "-a", "1.0", "-s", "8", "--seed", "4", "--backend", "cuda", "--gpus", str(g),
                            "--out-dir", str(out), "cornell-box"], stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        assert f"{g} GPU(s)" in r.stderr
        outs.append((out / "image_0000.png").read_bytes())
    assert outs[0] == outs[1]
