"""Host front end (C++ mirror of the reference's scene API): scenes, flatten, OBJ loader, console_app."""
import os
import subprocess

import numpy as np
import pytest

import raytracer_weekend_b200 as rtw
from conftest import ROOT, bits

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


# ---- ProgressMessage stream (lib.rs:128-138) in the host receivers' wire format: postcard 0.7.3 + COBS --------------
def _cobs_decode(frame: bytes) -> bytes:
    """Independent COBS decoder (Cheshire & Baker); `frame` without its 0x00 terminator."""
    out, i = bytearray(), 0
    while i < len(frame):
        code = frame[i]
        assert code != 0
        out += frame[i + 1:i + code]
        i += code
        if code != 0xFF and i < len(frame):
            out.append(0)
    return bytes(out)


def test_progress_messages_known_answers():
    import ctypes as C
    h = rtw.host_lib()
    buf = (C.c_uint8 * 64)()
    # ImageEnd: variant 2, no payload -> raw 02 -> COBS 02 02, terminator 00
    n = h.rtwh_progress_image_end(buf, 64)
    assert bytes(buf[:n]) == b"\x02\x02\x00"
    # ImageStart{32, 32, 50} (discovery_app/src/bin/raytracer.rs:55-66): raw = 00 | 20 00 00 00 | 20 00 00 00 | 32 00 00 00;
    # COBS by hand: zero-separated groups "", "20", "", "", "20", "", "", "32", "", "", "" -> code = len + 1 each
    n = h.rtwh_progress_image_start(32, 32, 50, buf, 64)
    assert bytes(buf[:n]) == bytes.fromhex("01 02 20 01 01 02 20 01 01 02 32 01 01 01 00")
    # Pixel{row 1, column 2, color (1.0, 0.5, -2.0)}: raw = 01 | 01 00 00 00 | 02 00 00 00 | 00 00 80 3f | 00 00 00 3f | 00 00 00 c0
    col = (C.c_float * 3)(1.0, 0.5, -2.0)
    n = h.rtwh_progress_pixel(1, 2, col, buf, 64)
    raw = bytes.fromhex("01 01000000 02000000 0000803f 0000003f 000000c0")
    assert buf[n - 1] == 0 and 0 not in bytes(buf[:n - 1])
    assert _cobs_decode(bytes(buf[:n - 1])) == raw
    assert bytes(buf[:n]) == bytes.fromhex("03 01 01 01 01 02 02 01 01 01 01 03 80 3f 01 01 02 3f 01 01 02 c0 00")
    assert h.rtwh_progress_pixel(1, 2, col, buf, 8) == rtw.RTW_ERR_INVALID     # buffer too small


def test_progress_frame_stream_decodes_to_the_pixel_order_of_the_reference():
    import struct
    rs = np.random.RandomState(5)
    hgt, wid, spp = 5, 7, 9
    accum = rs.uniform(0, 4, (hgt, wid, 3)).astype(np.float32)
    accum[0, 0] = 0.0                       # zeros inside the payload exercise the byte stuffing
    stream = rtw.progress_frame(accum, spp)
    frames = stream.split(b"\x00")
    assert frames[-1] == b"" and len(frames) == 1 + hgt * wid + 1 + 1
    msgs = [_cobs_decode(f) for f in frames[:-1]]
    assert msgs[0] == bytes([0]) + struct.pack("<III", wid, hgt, spp)
    assert msgs[-1] == bytes([2])
    k = 1
    for y_top in range(hgt):                # (0..h).rev() x (0..w): top image row first, Pixel.row counts from the bottom
        for col in range(wid):
            tag, row, column, r, g, b = struct.unpack("<BIIfff", msgs[k])
            assert (tag, row, column) == (1, hgt - 1 - y_top, col)
            assert np.array_equal(np.float32([r, g, b]), accum[y_top, col])
            k += 1


def test_cobs_long_runs_without_zero():
    # a 300-byte zero-free payload needs the 0xFF block split; no message of the protocol is that long, so the
    # encoder is exercised through a frame whose pixel colours are chosen zero-free and checked by the decoder
    accum = np.full((2, 200, 3), np.float32(1.2345678), np.float32)
    stream = rtw.progress_frame(accum, 1)
    frames = stream.split(b"\x00")[:-1]
    assert len(frames) == 402
    for f in frames[1:-1]:
        raw = _cobs_decode(f)
        assert len(raw) == 21 and raw[0] == 1


def test_oracle_render_frames_equals_frame_by_frame(oracle):
    with rtw.Scene.from_name(oracle, "cornell-box", 1.0, seed=1) as s:
        cam = s.cameras[0]
        cam2 = rtw.camera_new((278, 278, -700), (278, 278, 0), (0, 1, 0), 40.0, 1.0)
        p = s.params(24, 24, 3, seed=11)
        got = {}

        def on_frame(i, accum, st):
            got[i] = (accum, st["segments"])
            return i < 1                    # stop after the second frame

        n = s.render_frames([cam, cam2, cam], p, on_frame)
        assert n == 2 and sorted(got) == [0, 1]
        for i, c in enumerate([cam, cam2]):
            a, st = s.render(c, s.params(24, 24, 3, seed=11 + i))
            assert np.array_equal(a, got[i][0]) and st.segments == got[i][1]


# ---- OBJ ingest fast path: the parallel parser must reproduce the reference parser exactly -------
TRICKY_OBJ = """# comment line, CRLF endings, signs, exponents, negative indices, quads, mixed corner forms\r
mtllib lib.mtl\r
v 0 0 0\r
v 1.5 +2.25 -3e-1
v 1 1 0   # trailing comment tokens are ignored by `>>`-style parsing of three numbers
v 0 1 0
v 0.1 0.2
vt 0 0
vt 1 0
vt 1 1
vn 0 0 1
vn 0 1 0
g group1
usemtl first
f 1/1/1 2/2/1 3/3/1 4/1/1
f -5 -4 -3
usemtl second
f 1//2 2//2 3//2
s off
f 1/1 2/2 3/3 4/3 5/1
usemtl
f 3 2 1
v 9 9 9
f -1 1 2
"""


def test_parallel_obj_parser_equals_the_reference_parser(tmp_path):
    obj = tmp_path / "tricky.obj"
    obj.write_bytes(TRICKY_OBJ.encode())
    ref = rtw.parse_obj(str(obj), mode=1)
    assert ref[0] == 2 + 1 + 1 + 3 + 1 + 1                      # quads / pentagon fan-triangulated
    for threads in (1, 2, 3, 8):
        got = rtw.parse_obj(str(obj), mode=0, threads=threads)
        assert got[:2] == ref[:2], threads
    # a file large enough to be cut into many chunks (>= 64 KiB each): vertices first, faces after, materials changing
    rs = np.random.RandomState(3)
    nv, nf = 20000, 60000
    lines = ["mtllib x.mtl\n"]
    lines += ["v %.6f %.6f %.6f\n" % tuple(r) for r in rs.uniform(-50, 50, (nv, 3))]
    lines += ["vn %.4f %.4f %.4f\n" % tuple(r) for r in rs.uniform(-1, 1, (nv, 3))]
    lines += ["vt %.5f %.5f\n" % tuple(r) for r in rs.uniform(0, 1, (nv, 2))]
    for i, (a, b, c, d) in enumerate(rs.randint(1, nv + 1, (nf, 4))):
        if i % 7000 == 0:
            lines.append("usemtl m%d\n" % (i // 7000 % 3))
        kind = i % 4
        if kind == 0:
            lines.append("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, c, c, c))
        elif kind == 1:
            lines.append("f %d//%d %d//%d %d//%d %d//%d\n" % (a, a, b, b, c, c, d, d))
        elif kind == 2:
            lines.append("f %d %d %d\n" % (a - nv - 1, b - nv - 1, c - nv - 1))      # relative indices
        else:
            lines.append("f %d/%d %d/%d %d/%d\n" % (a, a, b, b, c, c))
    big = tmp_path / "big.obj"
    big.write_text("".join(lines))
    ref = rtw.parse_obj(str(big), mode=1)
    assert ref[0] == nf + nf // 4
    for threads in (1, 4, 0):
        assert rtw.parse_obj(str(big), mode=0, threads=threads)[:2] == ref[:2], threads


def test_parallel_obj_parser_reports_the_same_errors(tmp_path):
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nf 1 2\n")
    for mode in (0, 1):
        with pytest.raises(rtw.RtwError, match="points / lines"):
            rtw.parse_obj(str(bad), mode=mode)
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n")
    for mode in (0, 1):
        with pytest.raises(rtw.RtwError, match="index out of range"):
            rtw.parse_obj(str(bad), mode=mode)
    bad.write_text("v 0 0 0\nv 1 0 0\nf 1 2 3\nv 0 1 0\n")         # forward reference: vertex 3 is defined after the face
    for mode in (0, 1):
        with pytest.raises(rtw.RtwError, match="index out of range"):
            rtw.parse_obj(str(bad), mode=mode)
    with pytest.raises(rtw.RtwError, match="cannot open"):
        rtw.parse_obj(str(tmp_path / "missing.obj"))


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference models are only present in the build container")
@pytest.mark.parametrize("stem", ["cow-nonormals", "monument_downscaled_polygon_reduced", "capsule", "Normals_Try3"])
def test_parallel_obj_parser_on_the_reference_models(stem):
    ref = rtw.parse_obj(f"{REF_MODELS}/{stem}.obj", mode=1)
    for threads in (1, 0):
        assert rtw.parse_obj(f"{REF_MODELS}/{stem}.obj", mode=0, threads=threads)[:2] == ref[:2]


# ---- PNG textures (image_texture.rs:23-30): lossless, so the decoder must reproduce the pixels exactly ---------------
def _png_bytes(img: np.ndarray, color_type: int, depth: int = 8, palette=None, filters=(0, 1, 2, 3, 4)) -> bytes:
    """A PNG written by hand so that every scanline filter type occurs (PIL picks filters by heuristics)."""
    import struct
    import zlib

    h, w = img.shape[:2]
    rows = img.reshape(h, -1).astype(np.uint8)
    if depth < 8:                                   # pack `depth`-bit samples MSB first
        per = 8 // depth
        padded = np.zeros((h, (w + per - 1) // per * per), np.uint8)
        padded[:, :w] = rows
        rows = sum(padded[:, k::per].astype(np.uint16) << (8 - depth * (k + 1)) for k in range(per)).astype(np.uint8)
    bpp = max(1, rows.shape[1] * 8 // w // 8) if depth == 8 else 1
    out, prev = bytearray(), np.zeros(rows.shape[1], np.int32)
    for y in range(h):
        cur = rows[y].astype(np.int32)
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        ft = filters[y % len(filters)]
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (a + prev) // 2
        else:
            p = a + prev - c
            pa, pb, pc = abs(p - a), abs(p - prev), abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 0))
    if palette is not None:
        png += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    data = zlib.compress(bytes(out), 6)
    return png + chunk(b"IDAT", data[:len(data) // 2]) + chunk(b"IDAT", data[len(data) // 2:]) + chunk(b"IEND", b"")


def test_png_decoder_reproduces_every_supported_layout(tmp_path):
    rs = np.random.RandomState(11)
    w, h = 37, 23
    rgb = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    rgba = np.concatenate([rgb, rs.randint(0, 256, (h, w, 1)).astype(np.uint8)], axis=2)
    grey = rs.randint(0, 256, (h, w)).astype(np.uint8)
    ga = np.stack([grey, rs.randint(0, 256, (h, w)).astype(np.uint8)], axis=2)
    pal = rs.randint(0, 256, (200, 3)).astype(np.uint8)
    idx = rs.randint(0, 200, (h, w)).astype(np.uint8)
    cases = {
        "rgb": (_png_bytes(rgb, 2), rgb),
        "rgba": (_png_bytes(rgba, 6), rgb),                                  # alpha dropped (image_texture.rs:44-50)
        "grey": (_png_bytes(grey, 0), np.repeat(grey[:, :, None], 3, 2)),
        "grey_alpha": (_png_bytes(ga, 4), np.repeat(grey[:, :, None], 3, 2)),
        "palette": (_png_bytes(idx, 3, palette=pal), pal[idx]),
        "grey4": (_png_bytes(grey >> 4, 0, depth=4), np.repeat(((grey >> 4) * 17)[:, :, None], 3, 2)),
        "grey1": (_png_bytes(grey >> 7, 0, depth=1), np.repeat(((grey >> 7) * 255)[:, :, None], 3, 2)),
        "palette2": (_png_bytes(idx & 3, 3, depth=2, palette=pal[:4]), pal[idx & 3]),
    }
    for name, (data, want) in cases.items():
        p = tmp_path / f"{name}.png"
        p.write_bytes(data)
        got = rtw.open_image(str(p))
        assert got.shape == want.shape and np.array_equal(got, want), name
    # files written by an independent encoder, decoded by an independent decoder
    from PIL import Image
    for mode, arr in (("RGB", rgb), ("RGBA", rgba), ("L", grey), ("P", idx)):
        im = Image.fromarray(arr, mode)
        if mode == "P":
            im.putpalette(pal.tobytes())
        p = tmp_path / f"pil_{mode}.png"
        im.save(p, optimize=(mode == "RGB"))
        assert np.array_equal(rtw.open_image(str(p)), np.asarray(Image.open(p).convert("RGB"))), mode


def test_png_decoder_refuses_what_it_cannot_reproduce(tmp_path):
    import struct
    rgb = np.zeros((4, 4, 3), np.uint8)
    good = _png_bytes(rgb, 2)
    p = tmp_path / "x.png"
    p.write_bytes(good[:-20])
    with pytest.raises(rtw.RtwError, match="truncated|CRC|IDAT"):
        rtw.open_image(str(p))
    bad = bytearray(good)
    bad[40] ^= 0xFF                                                          # corrupt the IDAT payload: CRC check
    p.write_bytes(bytes(bad))
    with pytest.raises(rtw.RtwError, match="CRC"):
        rtw.open_image(str(p))
    ihdr16 = good.replace(struct.pack(">IIBBBBB", 4, 4, 8, 2, 0, 0, 0), struct.pack(">IIBBBBB", 4, 4, 16, 2, 0, 0, 0))
    p.write_bytes(ihdr16)
    with pytest.raises(rtw.RtwError, match="CRC|bit depth"):                 # (the IHDR CRC no longer matches either)
        rtw.open_image(str(p))
    from PIL import Image
    Image.fromarray(np.arange(64, dtype=np.uint16).reshape(8, 8) * 900).save(tmp_path / "g16.png")      # 16-bit grey
    with pytest.raises(rtw.RtwError, match="bit depth 16"):
        rtw.open_image(str(tmp_path / "g16.png"))
    with pytest.raises(rtw.RtwError, match="file not found"):              # JPEG stays a pre-decoded asset
        rtw.open_image(str(tmp_path / "nothing.jpg"))


def test_progress_stream_through_the_receiver_stand_in():
    """tools/progress_receiver.py restates discovery_host_receiver/src/main.rs:56-101 (tonemap per pixel, put_pixel(column,
    row), rotate180 at ImageEnd).  Pixel.row counts from the bottom (lib.rs:58), so after the rotation the receiver's
    image is the frame mirrored left-right — the same thing it shows for the embedded renderer."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("progress_receiver", os.path.join(ROOT, "tools", "progress_receiver.py"))
    recv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(recv)
    rs = np.random.RandomState(8)
    h, w, spp = 6, 9, 4
    accum = rs.uniform(0, 6, (h, w, 3)).astype(np.float32)
    frames = list(recv.receive(rtw.progress_frame(accum, spp) + rtw.progress_frame(accum * 0.5, spp)))
    assert len(frames) == 2 and frames[0].shape == (h, w, 3)
    c = np.clip(np.sqrt(accum * np.float32(1.0 / spp)), 0, np.float32(0.999))
    top_down = (np.float32(255.999) * c).astype(np.uint8)        # the frame as console_app writes it (main.rs:66-90)
    # receiver: y = Pixel.row = h-1-y_top, then rotate180 -> y_top, x mirrored
    assert np.array_equal(frames[0], top_down[:, ::-1])


def test_world_generated_once_flattens_like_from_name(oracle):
    """rtwh_world_create + rtwh_world_flatten = the two halves of rtwh_build_scene: same canonical ids, same cameras,
    same closest hits, and one World can feed several scenes."""
    with rtw.World("jumpy-balls", 16 / 9, seed=4) as world:
        assert len(world.cameras) == 1 and world.background == pytest.approx((0.7, 0.8, 1.0))
        with rtw.Scene.from_world(oracle, world) as a, rtw.Scene.from_world(oracle, world) as b, \
                rtw.Scene.from_name(oracle, "jumpy-balls", 16 / 9, seed=4) as c:
            assert a.num_prims == b.num_prims == c.num_prims
            assert bytes(a.cameras[0]) == bytes(c.cameras[0])
            rays = oracle.capture_rays(c, c.cameras[0], 64, 36, 3, 0, 0)
            ha, hb, hc = a.trace_closest(rays), b.trace_closest(rays), c.trace_closest(rays)
            assert np.array_equal(ha["prim_id"], hc["prim_id"]) and np.array_equal(hb["t"], hc["t"])
            assert np.array_equal(ha["material_id"], hc["material_id"]) and np.array_equal(bits(ha["normal"]), bits(hc["normal"]))
    with pytest.raises(rtw.RtwError, match="unknown scene"):
        rtw.World("nope", 1.0)


# ---- JPEG ingest (image_texture.rs:23-30; host/jpeg_reader.cpp) ---------------------------------------------------------
def test_jpeg_decoder_reproduces_the_committed_earthmap_fixture():
    """VERDICT r01 missing #6: ImageTexture::open("models/earthmap.jpg") from the FILE.  The decoder follows libjpeg's
    default arithmetic (islow IDCT, fixed-point YCbCr -> RGB), so its texels equal the committed PIL-decoded fixture
    assets/earthmap.rtwi bit for bit — the stated +-1 LSB decoder caveat only concerns zune-jpeg."""
    import os
    jpg = os.path.join(rtw.ASSET_DIR, "models", "earthmap.jpg")
    got = rtw.open_image(jpg)
    want = rtw.read_rtwi(os.path.join(rtw.ASSET_DIR, "earthmap.rtwi"))
    assert got.shape == (512, 1024, 3) and np.array_equal(got, want)
    from PIL import Image
    assert np.array_equal(got, np.asarray(Image.open(jpg).convert("RGB")))
    # the scenes' own path resolves to the JPEG file under the asset directory
    assert np.array_equal(rtw.open_image("models/earthmap.jpg"), want)


def test_jpeg_decoder_against_pil_on_synthetic_files(tmp_path):
    """baseline and progressive, 4:4:4 / 4:2:2 / 4:2:0 (libjpeg's fancy upsampling), restart markers, optimised Huffman
    tables, grey, sizes that are not multiples of the MCU: every texel equal to PIL's (libjpeg-turbo)."""
    from PIL import Image
    rs = np.random.RandomState(0)

    def img(w, h):
        y, x = np.mgrid[0:h, 0:w]
        base = np.stack([128 + 100 * np.sin(x / 7.0) * np.cos(y / 9.0), 128 + 90 * np.cos(x / 5.0 + y / 11.0), (x * 3 + y * 5) % 256], -1)
        return np.clip(base + rs.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)

    path = str(tmp_path / "t.jpg")
    cases = 0
    for (w, h) in [(64, 48), (37, 29), (1, 1), (17, 8), (200, 133)]:
        for kw in [dict(quality=90, subsampling=0), dict(quality=75, subsampling=1), dict(quality=60, subsampling=2),
                   dict(quality=85, subsampling=2, progressive=True), dict(quality=95, subsampling=0, progressive=True),
                   dict(quality=80, subsampling=2, restart_marker_blocks=3), dict(quality=50, subsampling=0, optimize=True)]:
            Image.fromarray(img(w, h)).save(path, "JPEG", **kw)
            assert np.array_equal(rtw.open_image(path), np.asarray(Image.open(path).convert("RGB"))), (w, h, kw)
            cases += 1
        Image.fromarray(img(w, h)[..., 0]).save(path, "JPEG", quality=80)
        assert np.array_equal(rtw.open_image(path), np.asarray(Image.open(path).convert("RGB"))), ("grey", w, h)
    assert cases == 35


def test_jpeg_decoder_errors(tmp_path):
    p = tmp_path / "bad.jpg"
    p.write_bytes(b"\xff\xd8\xff\xc9\x00\x0b\x08\x00\x10\x00\x10\x01\x01\x11\x00")      # SOF9: arithmetic coding
    with pytest.raises(rtw.RtwError, match="unsupported coding process"):
        rtw.open_image(str(p))
    p.write_bytes(b"\xff\xd8\xff\xd9")
    with pytest.raises(rtw.RtwError, match="no image data"):
        rtw.open_image(str(p))
    with pytest.raises(rtw.RtwError, match="not found"):
        rtw.open_image(str(tmp_path / "missing.jpg"))


def test_scene_constants_match_scenes_rs():
    """host/scenes.cpp is a transcription of console_app/src/scenes.rs, and BOTH the oracle and the CUDA path are fed by it: a
    slip in a scene constant would be invisible to every parity test.  tests/golden/scenes_rs_literals.json holds the
    floating-point literals of every scene function of the reference (tools/make_scene_literals.py); the mirror must use
    exactly the same multiset of literals per scene (helpers expanded; the BASELINE variants of the cow / monument scenes may
    add literals, never drop one)."""
    import collections
    import json
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "scenes_rs_literals.json")))["functions"]
    src = re.sub(r"//[^\n]*", "", open(os.path.join(root, "raytracer-weekend_b200", "host", "scenes.cpp")).read())
    lit = re.compile(r"(?<![\w.])-?\d+\.\d+(?=f?\b)")
    helpers = {"ground_checker()": [0.2, 0.3, 0.1, 0.9, 0.9, 0.9, 10.0],   # scenes.rs writes the checker out in every scene
               "make_cam(": [0.0, 1.0]}                                     # ... and time0 / time1 of Camera::new
    bodies = {}
    for m in re.finditer(r"^World (\w+)\(", src, re.M):
        end = src.find("\n}\n", m.start())
        bodies[m.group(1)] = src[m.start():end]
    names = {"wavefront_cow_obj": "wavefront_cow", "wavefront_suspension_obj": None}
    checked = 0
    for fn, ref_lits in ref.items():
        mine = names.get(fn, fn)
        if mine is None:
            continue   # loaded through the generic OBJ path: no function of its own in the mirror
        assert mine in bodies, f"scene {fn} has no mirror"
        body = bodies[mine]
        got = [float(x) for x in lit.findall(body)]
        for call, extra in helpers.items():
            got += extra * body.count(call)
        want, have = collections.Counter(ref_lits), collections.Counter(got)
        if mine in ("wavefront_cow", "textured_monument"):   # + the BASELINE.json variant (Lambertian + metal / earth sphere)
            assert not (want - have), (fn, "missing", dict(want - have))
        else:
            assert want == have, (fn, "missing", dict(want - have), "extra", dict(have - want))
        checked += 1
    assert checked >= 11
