"""Host front end (C++ mirror of the reference's scene API): scenes, flatten, OBJ loader, console_app."""
import os
import subprocess

import numpy as np
import pytest

import raytracer_weekend_b200 as rtw
from conftest import ROOT

REF_MODELS = "/root/reference/models"


def test_scene_list_matches_the_reference_subcommands():
    names = rtw.scene_names()
    for n in ["jumpy-balls", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell-box",
              "smokey-cornell-box", "book2-final-scene", "animated-book2-final-scene", "simple-triangle",
              "wavefront-cow-obj", "wavefront-suspension-obj", "textured-monument"]:  # scenes.rs:24-39
        assert n in names


@pytest.mark.parametrize("name,prims", [("cornell-box", 18), ("two-spheres", 2), ("two-perlin-spheres", 2), ("earth", 1),
                                        ("simple-light", 4), ("simple-triangle", 2), ("wavefront-cow-obj", 5806),
                                        ("cow-lambert-metal", 5806), ("monument-earth", 7800),
                                        ("stress:300:20", 500), ("smokey-cornell-box", 8), ("book2-final-scene", 3409)])
def test_scenes_flatten_into_a_sink(oracle, name, prims):
    with rtw.Scene.from_name(oracle, name, 16 / 9, seed=3) as s:
        assert s.num_prims == prims
        assert len(s.cameras) == 1


def test_jumpy_balls_is_seeded_and_faithful(oracle):
    with rtw.Scene.from_name(oracle, "jumpy-balls", 16 / 9, seed=1) as a, \
            rtw.Scene.from_name(oracle, "jumpy-balls", 16 / 9, seed=1) as b, \
            rtw.Scene.from_name(oracle, "jumpy-balls", 16 / 9, seed=2) as c:
        # 5 fixed spheres + up to 22*22 moving ones minus those near (4, 0.2, 0) (scenes.rs:74-140)
        assert 5 + 400 < a.num_prims <= 5 + 484
        assert a.num_prims == b.num_prims
        rays = rtw.make_rays(np.tile([[13, 2, 3]], (64, 1)), np.random.RandomState(0).uniform(-1, 0, (64, 3)) * [1, .2, .3])
        assert np.array_equal(a.trace_closest(rays)["t"], b.trace_closest(rays)["t"])
        assert not np.array_equal(a.trace_closest(rays)["t"], c.trace_closest(rays)["t"])
        assert a.background == pytest.approx((0.7, 0.8, 1.0))
        assert a.cameras[0].lens_radius == pytest.approx(0.05)   # aperture 0.1 (scenes.rs:147)


def test_unsupported_scenes_fail_with_a_reason(oracle):
    with pytest.raises(rtw.RtwError, match="usemtl without mtllib"):
        rtw.Scene.from_name(oracle, "wavefront-suspension-obj", 1.0)
    with pytest.raises(rtw.RtwError, match="no decoded image"):   # the PNG is missing from the reference tree too (.MISSING_LARGE_BLOBS)
        rtw.Scene.from_name(oracle, "textured-monument", 1.0)
    with pytest.raises(rtw.RtwError, match="unknown scene"):
        rtw.Scene.from_name(oracle, "nope", 1.0)


def test_animated_scene_has_thirty_cameras(oracle):
    with rtw.Scene.from_name(oracle, "animated-book2-final-scene", 1.0, seed=1) as s:   # scenes.rs:622-667
        assert len(s.cameras) == 30 and s.num_prims == 3409
        assert s.cameras[0].origin[0] == 478.0 and s.cameras[0].lens_radius == 0.5
        assert abs(s.cameras[29].origin[0] - (478.0 - 29 * 2 * 478.0 / 30)) < 1e-3


def test_obj_loader_semantics(tmp_path):
    obj = tmp_path / "t.obj"
    obj.write_text("""# quad + triangle, negative indices, mixed normals
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vn 0 0 1
f 1/1/1 2/2/1 3/3/1 4/1/1
f -4 -3 -2
""")
    v, n, uv = rtw.load_obj(str(obj))
    assert v.shape == (3, 9)                                   # quad fan-triangulated: (1,2,3), (1,3,4)
    np.testing.assert_array_equal(v[0], [0, 0, 0, 1, 0, 0, 1, 1, 0])
    np.testing.assert_array_equal(v[1], [0, 0, 0, 1, 1, 0, 0, 1, 0])
    np.testing.assert_array_equal(v[2], [0, 0, 0, 1, 0, 0, 1, 1, 0])
    np.testing.assert_array_equal(n[0], [0, 0, 1] * 3)
    np.testing.assert_array_equal(n[2], [0, 0, 1] * 3)          # missing normals -> face normal (b-a)x(c-a) (triangular.rs:53-55)
    np.testing.assert_array_equal(uv[0], [0, 0, 1, 0, 1, 1])
    np.testing.assert_array_equal(uv[2], [0, 0, 1, 0, 0, 1])    # missing uvs -> defaults (triangular.rs:57-65)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nl 1 2\nf 1 2\n")
    with pytest.raises(rtw.RtwError, match="points / lines"):
        rtw.load_obj(str(bad))


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference models are only present in the build container")
@pytest.mark.parametrize("stem", ["cow-nonormals", "monument_downscaled_polygon_reduced"])
def test_obj_loader_reproduces_the_committed_fixtures(stem):
    v, n, uv = rtw.load_obj(f"{REF_MODELS}/{stem}.obj")
    fv, fn, fuv = rtw.read_rtwm(os.path.join(rtw.ASSET_DIR, stem + ".rtwm"))
    assert np.array_equal(v, fv)
    assert (n is None) == (fn is None) and (n is None or np.array_equal(n, fn))
    assert (uv is None) == (fuv is None) and (uv is None or np.array_equal(uv, fuv))


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference models are only present in the build container")
def test_earthmap_fixture_is_the_decoded_reference_image():
    from PIL import Image

    ref = np.asarray(Image.open(f"{REF_MODELS}/earthmap.jpg").convert("RGB"))
    assert np.array_equal(ref, rtw.read_rtwi(os.path.join(rtw.ASSET_DIR, "earthmap.rtwi")))


def test_console_app_cli():
    exe = os.path.join(rtw.PKG_DIR, "bin", "console_app")
    r = subprocess.run([exe, "--backend", "cpu", "cornell-box"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "not available" in r.stderr
    r = subprocess.run([exe], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "missing scene" in r.stderr and "jumpy-balls" in r.stderr
    r = subprocess.run([exe, "-w", "8", "nope"], stderr=subprocess.PIPE, text=True, cwd=ROOT)
    assert r.returncode == 1 and "unknown scene" in r.stderr
