#!/usr/bin/env python3
"""Stand-in for discovery_host_receiver (src/main.rs:11-110) for checking a ProgressMessage stream without Rust:
reads postcard 0.7 + COBS frames (zero-delimited), rebuilds the image exactly like the receiver does — divide by
samples_per_pixel, sqrt, clamp to 0.999, scale by 255.999, `put_pixel(column, row)`, rotate by 180 degrees at
ImageEnd (main.rs:74-101) — and writes it as a PNG.

    raytracer-weekend_b200/bin/console_app -w 200 -a 1.0 -s 50 --progress-out stream.bin cornell-box
    tools/progress_receiver.py stream.bin foo.png

The reference's receiver is untouched by this repository; this script only demonstrates that the bytes the CUDA
back end streams are the ones it expects."""
import struct
import sys

import numpy as np


def cobs_decode(frame: bytes) -> bytes:
    out, i = bytearray(), 0
    while i < len(frame):
        code = frame[i]
        if code == 0:
            raise ValueError("zero inside a COBS frame")
        out += frame[i + 1:i + code]
        i += code
        if code != 0xFF and i < len(frame):
            out.append(0)
    return bytes(out)


def receive(stream: bytes):
    """Yield one uint8 image [h, w, 3] per ImageStart .. ImageEnd."""
    img = spp = None
    for frame in stream.split(b"\x00"):
        if not frame:
            continue
        msg = cobs_decode(frame)
        tag = msg[0]
        if tag == 0:                                            # ImageStart { width, height, samples_per_pixel }
            w, h, spp = struct.unpack("<III", msg[1:13])
            img = np.zeros((h, w, 3), np.uint8)
        elif tag == 1 and img is not None:                      # Pixel(Pixel { row, column, color })
            row, column, r, g, b = struct.unpack("<IIfff", msg[1:21])
            c = np.sqrt(np.float32([r, g, b]) * np.float32(1.0 / spp))
            c = np.clip(c, 0.0, np.float32(0.999))
            img[row, column] = (np.float32(255.999) * c).astype(np.uint8)       # get_pixel_mut(column, row): y = row
        elif tag == 2 and img is not None:                      # ImageEnd: rotate180, save
            yield img[::-1, ::-1].copy()
            img = None
        elif tag > 2:
            raise ValueError(f"unknown ProgressMessage variant {tag}")


def main():
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    from PIL import Image
    frames = list(receive(open(sys.argv[1], "rb").read()))
    for k, im in enumerate(frames):
        name = sys.argv[2] if len(frames) == 1 else sys.argv[2].replace(".png", f"_{k:04d}.png")
        Image.fromarray(im).save(name)
        print(f"frame {k}: {im.shape[1]}x{im.shape[0]} -> {name}")


if __name__ == "__main__":
    main()
