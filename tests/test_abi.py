"""The C-ABI boundary: every entry point include/*.h declares is exported, struct layouts match the
bindings, flattening is correct (host logic, no GPU needed) and — without a CUDA device — every
compute call fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import raytracer_weekend_b200 as rtw
from conftest import HAS_GPU, ROOT, ORACLE_LIB

HDR = os.path.join(ROOT, "include", "rtw_cuda.h")
SINK_HDR = os.path.join(ROOT, "include", "rtw_sink.h")


def declared_functions(path, prefix):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", txt)))


def exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    return {l.split()[-1] for l in out.splitlines() if " T " in l}


def test_cuda_library_exports_every_declared_symbol():
    names = declared_functions(HDR, "rtw_")
    assert len(names) >= 35
    missing = [n for n in names if n not in exported(rtw.CUDA_LIB)]
    assert not missing, missing


def test_host_library_exports_sink_api():
    names = declared_functions(SINK_HDR, "rtwh_")
    assert set(names) >= {"rtwh_sink_open", "rtwh_sink_close", "rtwh_last_error"}
    assert not [n for n in names if n not in exported(rtw.HOST_LIB)]


def test_oracle_mirrors_the_emit_and_render_abi():
    # the oracle is driven through the same sink table: same names, prefix orc_
    ex = exported(ORACLE_LIB)
    for n in rtw.api._SINK_FUNCS:
        assert "orc_" + n in ex, n
    assert "orc_trace_closest" in ex


def test_product_libraries_do_not_link_the_oracle():
    for lib in (rtw.CUDA_LIB, rtw.HOST_LIB, os.path.join(rtw.PKG_DIR, "bin", "console_app")):
        out = subprocess.run(["ldd", lib], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle" not in out
        assert "orc_" not in subprocess.run(["nm", "-D", lib], stdout=subprocess.PIPE, text=True).stdout
    # and no product source mentions the oracle directory
    for base, _, files in os.walk(rtw.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle/" not in txt, os.path.join(base, f)


def test_struct_layouts_match_the_header():
    # sizes the C compiler sees (static_asserts compiled on the fly)
    src = f'''#include "{HDR}"
    #include <stdio.h>
    int main() {{ printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rtw_ray), sizeof(rtw_hit), sizeof(rtw_camera),
      sizeof(rtw_render_params), sizeof(rtw_render_stats), sizeof(rtw_build_stats), sizeof(rtw_bvh_node)); return 0; }}'''
    exe = "/tmp/rtw_sizes"
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert sizes == [rtw.RAY_DTYPE.itemsize, rtw.HIT_DTYPE.itemsize, C.sizeof(rtw.Camera), C.sizeof(rtw.RenderParams),
                     C.sizeof(rtw.RenderStats), C.sizeof(rtw.BuildStats), rtw.BVH_NODE_DTYPE.itemsize]


def test_flatten_canonical_order_and_instances():
    b = rtw.cuda_backend()
    with b.new_scene() as s:
        white = s.lambertian_rgb(.73, .73, .73)
        light = s.diffuse_light_rgb(15, 15, 15)
        assert s.yz_rect(0, 555, 0, 555, 555, white) == 0
        assert s.xz_rect(213, 343, 227, 332, 554, light) == 1
        s.push_translation((265, 0, 295))
        s.push_rotation_y(15.0)
        assert s.cuboid((0, 0, 0), (165, 330, 165), white) == 2
        s.pop_transform()
        s.pop_transform()
        s.begin_group()
        assert s.sphere((0, 0, 0), 1.0, white) == 8
        assert s.triangles(np.arange(18, dtype=np.float32).reshape(2, 9), white) == 9
        s.end_group()
        assert s.num_prims == 11
        # types: 2 yz, 3 xz, 4 xy, 0 sphere, 5 triangle ; cuboid order XY,XY,XZ,XZ,YZ,YZ (rectangular.rs:177-234)
        assert [s.prim_info(i)[0] for i in range(11)] == [2, 3, 4, 4, 3, 3, 2, 2, 0, 5, 5]
        assert [s.prim_info(i)[1] for i in range(11)] == [0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0]
        assert [s.prim_info(i)[2] for i in range(11)] == [white, light] + [white] * 9
        ops = s.instance_ops(1)
        assert [k for k, _ in ops] == [0, 1]  # outermost first: Translation then YRotation
        assert ops[0][1] == (265.0, 0.0, 295.0)
        rad = np.float32(15.0) * np.float32(np.float32(np.pi) / np.float32(180.0))
        assert ops[1][1][0] == np.float32(np.sin(np.float64(rad))) or abs(ops[1][1][0] - np.sin(float(rad))) < 1e-7
        assert abs(ops[1][1][1] - np.cos(float(rad))) < 1e-7


def test_emit_errors():
    b = rtw.cuda_backend()
    with b.new_scene() as s:
        with pytest.raises(rtw.RtwError, match="bad"):
            s.sphere((0, 0, 0), 1.0, 0)                      # no such material
        with pytest.raises(rtw.RtwError, match="fuzz"):
            s.metal(.5, .5, .5, 1.5)                          # assert!(fuzz <= 1.0), material.rs:71
        with pytest.raises(rtw.RtwError, match="texture"):
            s.lambertian(7)
        with pytest.raises(rtw.RtwError, match="no open transform"):
            s.pop_transform()
        s.push_translation((1, 2, 3))
        with pytest.raises(rtw.RtwError, match="empty instance"):
            s.pop_transform()
        with pytest.raises(rtw.RtwError, match="unbalanced"):
            s.build()
    with b.new_scene() as s:
        with pytest.raises(rtw.RtwError, match="empty scene"):
            s.build()
        m = s.lambertian_rgb(.5, .5, .5)
        for _ in range(8):
            s.push_rotation_y(1.0)
        with pytest.raises(rtw.RtwError, match="nesting"):
            s.push_rotation_y(1.0)
        with pytest.raises(rtw.RtwError, match="not built"):
            s.trace_closest(rtw.make_rays([[0, 0, 0]], [[0, 0, 1]]))


def test_medium_emit_rules():
    b = rtw.cuda_backend()
    with b.new_scene() as s:
        white = s.texture_solid(1, 1, 1)
        m = s.lambertian(white)
        assert s.sphere((0, 0, 0), 1.0, m) == 0
        assert s.begin_medium(0.01, white) == 1            # ConstantMedium = ONE canonical primitive ...
        s.push_translation((1, 2, 3))
        s.push_rotation_y(15.0)
        s.cuboid((0, 0, 0), (1, 1, 1), m)                   # ... its boundary gets no ids
        with pytest.raises(rtw.RtwError, match="exactly one boundary"):
            s.sphere((0, 0, 0), 1.0, m)
        s.pop_transform()
        s.pop_transform()
        s.end_medium()
        assert s.num_prims == 2 and s.prim_info(1)[0] == 7 and s.prim_info(1)[1] == 1
        assert s.begin_medium(0.2, white) == 2
        with pytest.raises(rtw.RtwError, match="one sphere or one cuboid"):
            s.xy_rect(0, 1, 0, 1, 0, m)
        with pytest.raises(rtw.RtwError, match="exactly one boundary"):
            s.end_medium()
        s.sphere((5, 5, 5), 2.0, m)
        s.end_medium()
        assert s.prim_info(2)[0] == 6
        s.push_translation((1, 0, 0))
        with pytest.raises(rtw.RtwError, match="transformed medium"):
            s.begin_medium(0.1, white)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_fallback():
    b = rtw.cuda_backend()
    assert b.device_count() < 0 and "CUDA" in b.last_error()
    with b.new_scene() as s:
        s.sphere((0, 0, 0), 1.0, s.lambertian_rgb(.5, .5, .5))
        with pytest.raises(rtw.RtwError, match="no CPU fallback"):
            s.build()
    with pytest.raises(rtw.RtwError):
        rtw.Scene.from_name(b, "cornell-box", 1.0)


def test_rotation_from_stored_sin_cos_equals_rotation_from_angle(oracle):
    """rtw_push_rotation_y_sincos takes what YRotation stores (transformations.rs:51-56); with sin / cos computed like
    transformations.rs:60-63 the flattened instance and every hit are identical to rtw_push_rotation_y(angle)."""
    import math
    rays = rtw.make_rays(np.tile([[0.3, 0.2, -5]], (256, 1)), np.random.RandomState(1).uniform(-0.3, 0.3, (256, 3)) + [0, 0, 1])
    hits = []
    for use_sincos in (False, True):
        s = oracle.new_scene()
        m = s.lambertian_rgb(0.5, 0.5, 0.5)
        s.push_translation((0.1, 0.2, 0.3))
        if use_sincos:
            rad = np.float32(15.0) * np.float32(np.float32(math.pi) / np.float32(180.0))
            s.push_rotation_y_sincos(float(np.sin(rad, dtype=np.float32)), float(np.cos(rad, dtype=np.float32)))
        else:
            s.push_rotation_y(15.0)
        s.cuboid((-1, -1, -1), (1, 1, 1), m)
        s.pop_transform()
        s.pop_transform()
        s.build()
        hits.append(s.trace_closest(rays))
        s.close()
    assert (hits[0]["prim_id"] >= 0).sum() > 50
    np.testing.assert_allclose(hits[0]["t"], hits[1]["t"], rtol=2e-6)      # numpy's sinf may differ from glibc's by an ulp
    assert np.array_equal(hits[0]["prim_id"], hits[1]["prim_id"])


def test_bulk_triangle_ingest_and_unbuilt_scene_errors():
    """rtw_add_triangles fills the scene arrays with all host threads from 2^16 triangles on: ids, types, instances and
    per-face materials must be what the sequential path gives; render entry points refuse an unbuilt scene."""
    b = rtw.cuda_backend()
    rs = np.random.RandomState(2)
    n = 70_000
    verts = rs.uniform(-1, 1, (n, 9)).astype(np.float32)
    with b.new_scene() as s:
        mats = [s.lambertian_rgb(.1 * k, .2, .3) for k in range(5)]
        ball = s.sphere((0, 0, 0), 1.0, mats[0])
        s.push_translation((1, 2, 3))
        mids = rs.randint(0, 5, n).astype(np.int32)
        first = s.triangles(verts, mats[0], material_ids=np.asarray(mats, np.int32)[mids])
        s.pop_transform()
        small = s.triangles(verts[:10], mats[3])
        assert (ball, first, small) == (0, 1, 1 + n) and s.num_prims == 1 + n + 10
        for i in (0, 1, 12345, n - 1):
            t, inst, m = s.prim_info(first + i)
            assert (t, inst, m) == (5, 1, mats[mids[i]])          # triangle, inside the translation, its own material
        assert s.prim_info(small + 9) == (5, 0, mats[3])
        cam = rtw.camera_new((0, 0, -5), (0, 0, 0), (0, 1, 0), 40.0, 1.0)
        with pytest.raises(rtw.RtwError, match="not built"):
            s.render(cam, s.params(8, 8, 1))
        with pytest.raises(rtw.RtwError, match="not built"):
            s.render_frames([cam], s.params(8, 8, 1), lambda i, a, st: True)


# ---- the Rust -sys crate against the C header (no Rust toolchain here: parse both) ---------------------------------------
_C2RUST = {"int": "c_int", "void": "c_void", "char": "c_char", "float": "f32", "uint8_t": "u8", "int32_t": "i32", "uint32_t": "u32",
           "uint64_t": "u64", "size_t": "usize"}
_RUST_SIZE = {"f32": 4, "u32": 4, "i32": 4, "u64": 8, "i64": 8, "u8": 1, "c_int": 4}


def _c_type_to_rust(t):
    """`const rtw_ray *` -> `*const rtw_ray`, `float[3]`-style parameters are pointers."""
    t = t.strip()
    stars = t.count("*")
    const = "const" in t.split("*")[0].split()
    base = [w for w in t.replace("*", " ").split() if w not in ("const", "struct")][0]
    base = _C2RUST.get(base, base)
    for _ in range(stars):
        base = ("*const " if const else "*mut ") + base
        const = False if stars > 1 else const
    return base


def _header_functions():
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"\n\s*((?:const\s+)?\w+\s*\*?)\s*(rtw_\w+)\s*\(([^;{]*?)\)\s*;", txt):
        if "typedef" in ret:
            continue
        params = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",")]:
            if a in ("void", ""):
                continue
            m = re.match(r"(.*?)(\w+)\s*(\[\d*\])?$", a)          # type, name, optional array suffix
            ctype = m.group(1) + ("*" if m.group(3) else "")
            if "rtw_frame_callback" in ctype:
                params.append("rtw_frame_callback")
            else:
                params.append(_c_type_to_rust(ctype))
        out[name] = (_c_type_to_rust(ret), params)
    return out


def _rust_functions():
    txt = open(os.path.join(ROOT, "rust", "raytracer_weekend_cuda_sys", "src", "lib.rs")).read()
    ext = txt[txt.index('extern "C" {'):]
    out = {}
    for name, args, ret in re.findall(r"pub fn (rtw_\w+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        params = [a.split(":", 1)[1].strip() for a in args.replace("\n", " ").split(",") if ":" in a]
        out[name] = ((ret or "()").strip(), params)
    return out


def _header_structs():
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", txt, flags=re.S):
        fields = []
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            m = re.match(r"(\w+)\s+(.*)$", decl)
            ctype = _C2RUST[m.group(1)]
            for nm in [x.strip() for x in m.group(2).split(",")]:
                arr = re.match(r"(\w+)\[(\d+)\]", nm)
                fields.append((arr.group(1), f"[{ctype}; {arr.group(2)}]") if arr else (nm, ctype))
        out[name] = fields
    return out


def _rust_structs():
    txt = open(os.path.join(ROOT, "rust", "raytracer_weekend_cuda_sys", "src", "lib.rs")).read()
    out = {}
    for attrs, name, body in re.findall(r"((?:#\[[^\]]*\]\s*)+)pub struct (\w+) \{(.*?)\n\}", txt, flags=re.S):
        if "repr(C)" not in attrs:
            continue
        body = re.sub(r"///.*", "", body)
        out[name] = [(f.split(":")[0].replace("pub", "").strip(), f.split(":")[1].strip()) for f in body.split(",") if ":" in f]
    return out


def _layout(fields):
    off, align_max, offs = 0, 1, []
    for _, t in fields:
        m = re.match(r"\[(\w+); (\d+)\]", t)
        base, n = (m.group(1), int(m.group(2))) if m else (t, 1)
        sz = _RUST_SIZE[base]
        off = (off + sz - 1) // sz * sz
        offs.append(off)
        off += sz * n
        align_max = max(align_max, sz)
    return offs, (off + align_max - 1) // align_max * align_max


def test_rust_sys_crate_matches_the_header():
    """VERDICT r01 missing #7: rust/raytracer_weekend_cuda_sys/src/lib.rs is kept in sync with include/rtw_cuda.h by
    hand and cannot be compiled here.  Every declared entry point must exist on the Rust side with the same arity and
    the same argument / return types; every #[repr(C)] struct must have the header's fields in the header's order with
    the same types — hence the same size and offsets, which are also compared with what the C compiler computes."""
    hf, rf = _header_functions(), _rust_functions()
    assert len(hf) >= 44 and set(hf) == set(rf), (sorted(set(hf) - set(rf)), sorted(set(rf) - set(hf)))
    for name, (ret, params) in hf.items():
        rret, rparams = rf[name]
        assert rret == ret, (name, ret, rret)
        assert rparams == params, (name, params, rparams)
    hs, rs = _header_structs(), _rust_structs()
    assert set(hs) <= set(rs) and len(hs) >= 7, sorted(set(hs) - set(rs))
    for name, fields in hs.items():
        assert rs[name] == fields, (name, fields, rs[name])
    # sizes / offsets from the real C compiler
    names = sorted(hs)
    probes = "".join(f'printf("{n} %zu", sizeof({n}));' + "".join(f'printf(" %zu", offsetof({n}, {f}));' for f, _ in hs[n]) +
                     'printf("\\n");' for n in names)
    src = f'#include "{HDR}"\n#include <stdio.h>\n#include <stddef.h>\nint main() {{ {probes} return 0; }}'
    exe = "/tmp/rtw_layout"
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    for line in subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines():
        tok = line.split()
        offs, size = _layout(rs[tok[0]])
        assert int(tok[1]) == size and [int(x) for x in tok[2:]] == offs, (tok[0], tok[1:], size, offs)
    # the abi version constant
    rust_txt = open(os.path.join(ROOT, "rust", "raytracer_weekend_cuda_sys", "src", "lib.rs")).read()
    ver = int(re.search(r"#define RTW_ABI_VERSION (\d+)", open(HDR).read()).group(1))
    assert f"pub const RTW_ABI_VERSION: c_int = {ver};" in rust_txt
    assert rtw.cuda_backend().fn("abi_version")() == ver


def test_every_build_recipe_compiles_the_same_translation_units():
    """csrc/Makefile (what build() runs), tools/ab_build.sh (A/B variants) and the -sys crate's build.rs (what a Rust
    maintainer runs) must list the same .cu files — a unit missing from one of them links without rtw_mem.cu / rtw_multi.cu."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "raytracer-weekend_b200", "csrc", "Makefile")).read()
    srcs = set(re.search(r"^SRCS := (.*)$", mk, re.M).group(1).split())
    on_disk = {f for f in os.listdir(os.path.join(root, "raytracer-weekend_b200", "csrc")) if f.endswith(".cu")}
    assert srcs == on_disk
    ab = open(os.path.join(root, "tools", "ab_build.sh")).read()
    assert {u + ".cu" for u in re.search(r"for f in ([\w ]+); do", ab).group(1).split()} == srcs
    rs = open(os.path.join(root, "rust", "raytracer_weekend_cuda_sys", "build.rs")).read()
    assert {u + ".cu" for u in re.findall(r'"(rtw_\w+)"', re.search(r"let units = \[(.*?)\];", rs, re.S).group(1))} == srcs
    hdrs = set(re.search(r"^HDRS := (.*)$", mk, re.M).group(1).split())
    cuh = {f for f in os.listdir(os.path.join(root, "raytracer-weekend_b200", "csrc")) if f.endswith(".cuh")}
    assert cuh <= hdrs, "a header the kernels include is missing from the Makefile's dependency list"
