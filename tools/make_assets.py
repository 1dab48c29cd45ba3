#!/usr/bin/env python3
"""Generate the binary asset fixtures under assets/ from the reference's model files.

/root/reference does not exist on the GPU box, so the meshes and the earth map that
BASELINE.json's configs name are converted ONCE, here, into two trivial binary formats that both
the C++ host library and numpy read:

  .rtwm  mesh   : "RTWM" u32 version=1, u32 ntris, u32 flags (1 = normals, 2 = uvs, 4 = the OBJ names
                  materials with usemtl — the fixture does not carry them),
                  f32 verts[ntris*9], f32 normals[ntris*9] (if flag 1), f32 uvs[ntris*6] (if flag 2)
                  One record per face in FILE ORDER (= canonical primitive order, triangular.rs:170-218);
                  polygons are fan-triangulated like the wavefront_obj crate does.
  .rtwi  image  : "RTWI" u32 version=1, u32 width, u32 height, u8 rgb[height*width*3]  (row 0 = top)
                  decoded with PIL; the reference decodes with image 0.25.2 / zune-jpeg, which may
                  differ by +-1 LSB — both oracle and GPU read THIS buffer.

Usage: python tools/make_assets.py [/root/reference/models] [assets]
"""
import struct
import sys

import numpy as np


def parse_obj(path):
    v, vt, vn = [], [], []
    tris = []  # per triangle: 3 x (vi, ti or None, ni or None)
    uses_mtl = False
    with open(path) as f:
        for line in f:
            parts = line.split()
            if not parts or parts[0].startswith("#"):
                continue
            if parts[0] == "v":
                v.append([float(x) for x in parts[1:4]])
            elif parts[0] == "vt":
                vals = [float(x) for x in parts[1:3]]
                vt.append(vals + [0.0] * (2 - len(vals)))
            elif parts[0] == "vn":
                vn.append([float(x) for x in parts[1:4]])
            elif parts[0] == "usemtl":
                uses_mtl = True
            elif parts[0] == "f":
                corners = []
                for c in parts[1:]:
                    idx = c.split("/")
                    vi = int(idx[0])
                    ti = int(idx[1]) if len(idx) > 1 and idx[1] else None
                    ni = int(idx[2]) if len(idx) > 2 and idx[2] else None
                    vi = vi - 1 if vi > 0 else len(v) + vi
                    if ti is not None:
                        ti = ti - 1 if ti > 0 else len(vt) + ti
                    if ni is not None:
                        ni = ni - 1 if ni > 0 else len(vn) + ni
                    corners.append((vi, ti, ni))
                if len(corners) < 3:
                    raise ValueError("points / lines panic in the reference (triangular.rs:186-191)")
                for k in range(2, len(corners)):
                    tris.append((corners[0], corners[k - 1], corners[k]))
    return np.array(v, np.float64), np.array(vt, np.float64), np.array(vn, np.float64), tris, uses_mtl


def write_mesh(path, v, vt, vn, tris, uses_mtl):
    n = len(tris)
    has_n = all(c[2] is not None for t in tris for c in t)
    has_t = all(c[1] is not None for t in tris for c in t)
    verts = np.zeros((n, 9), np.float32)
    norms = np.zeros((n, 9), np.float32)
    uvs = np.zeros((n, 6), np.float32)
    for i, t in enumerate(tris):
        for k, (vi, ti, ni) in enumerate(t):
            verts[i, 3 * k:3 * k + 3] = v[vi].astype(np.float32)  # f64 parse -> `as f32` (triangular.rs:153-166)
            if has_n:
                norms[i, 3 * k:3 * k + 3] = vn[ni].astype(np.float32)
            if has_t:
                uvs[i, 2 * k:2 * k + 2] = vt[ti].astype(np.float32)
    flags = (1 if has_n else 0) | (2 if has_t else 0) | (4 if uses_mtl else 0)
    with open(path, "wb") as f:
        f.write(b"RTWM" + struct.pack("<III", 1, n, flags))
        f.write(verts.tobytes())
        if has_n:
            f.write(norms.tobytes())
        if has_t:
            f.write(uvs.tobytes())
    print(f"{path}: {n} triangles, normals={has_n}, uvs={has_t}")


def write_image(path, src):
    from PIL import Image

    im = Image.open(src).convert("RGB")
    a = np.asarray(im, np.uint8)
    with open(path, "wb") as f:
        f.write(b"RTWI" + struct.pack("<III", 1, a.shape[1], a.shape[0]))
        f.write(a.tobytes())
    print(f"{path}: {a.shape[1]}x{a.shape[0]}")


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/models"
    dst = sys.argv[2] if len(sys.argv) > 2 else "assets"
    for name, out in (("cow-nonormals.obj", "cow-nonormals.rtwm"),
                      ("monument_downscaled_polygon_reduced.obj", "monument_downscaled_polygon_reduced.rtwm")):
        write_mesh(f"{dst}/{out}", *parse_obj(f"{src}/{name}"))
    write_image(f"{dst}/earthmap.rtwi", f"{src}/earthmap.jpg")


if __name__ == "__main__":
    main()
