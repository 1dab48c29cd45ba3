#!/usr/bin/env python3
"""Fixture generator (run where /root/reference exists): the floating-point literals of every scene function of the reference's
console_app/src/scenes.rs, in source order, into tests/golden/scenes_rs_literals.json.  tests/test_host.py compares the C++ mirror
(host/scenes.cpp) with it, so a transcription slip in a scene constant — invisible to every parity test, because the oracle and the
CUDA path are fed by the same front end — fails a CPU test."""
import json
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/console_app/src/scenes.rs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes_rs_literals.json")
LIT = re.compile(r"(?<![\w.])-?\d+\.\d+")

def strip_comments(text):
    text = re.sub(r"//[^\n]*", "", text)
    return text.replace("..", " .. ")   # Rust ranges: 0.5..1.0

def main():
    text = strip_comments(open(SRC).read())
    out = {}
    for m in re.finditer(r"^pub fn (\w+)\(", text, re.M):
        name = m.group(1)
        end = text.find("\n}\n", m.start())
        body = text[m.start():end]
        out[name] = [float(x) for x in LIT.findall(body)]
    consts = {c.group(1): [float(x) for x in LIT.findall(c.group(2))]
              for c in re.finditer(r"^const (\w+): Color = ([^;]+);", text, re.M)}
    # ... and the distinctive literals of the library itself (everything but 0.0 / 0.5 / 1.0 / 2.0), per source file: the constants
    # an op-for-op restatement must carry (t_min 0.001, the 0.0001 paddings, near_zero's 1e-8, the Hermite 3.0, ...)
    import glob
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(SRC))), "raytracer_weekend_lib", "src")
    lib_lit = re.compile(r"(?<![\w.])-?\d+\.\d+(?:e-?\d+)?|(?<![\w.])\d+e-?\d+")
    library = {}
    for path in sorted(glob.glob(os.path.join(lib, "**", "*.rs"), recursive=True)):
        body = strip_comments(open(path).read()).split("#[cfg(test)]")[0]
        vals = sorted({abs(float(x)) for x in lib_lit.findall(body)} - {0.0, 0.5, 1.0, 2.0})
        if vals:
            library[os.path.relpath(path, lib)] = vals
    json.dump({"source": "console_app/src/scenes.rs", "functions": out, "constants": consts, "library_literals": library},
              open(OUT, "w"), indent=0)
    print(library)
    print({k: len(v) for k, v in out.items()}, consts)

if __name__ == "__main__":
    main()
