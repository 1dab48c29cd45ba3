"""The documents cite tests, files and environment switches by name; this keeps the citations alive."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "INTEGRATION.md", "BASELINE.md", "README.md", "profiles/README.md", "tools/README.md"]


def _text(name):
    with open(os.path.join(ROOT, name)) as f:
        return f.read()


def _all_tests():
    names = set()
    for p in glob.glob(os.path.join(ROOT, "tests", "test_*.py")):
        names |= set(re.findall(r"^def (test_\w+)", open(p).read(), re.M))
    return names


def test_every_cited_test_exists():
    tests = _all_tests()
    missing = []
    for doc in DOCS:
        for tok in set(re.findall(r"`(?:tests/\w+\.py::)?(test_\w+?)(\*?)`", _text(doc))):
            name, star = tok
            if name.endswith("_"):
                star = "*"
            if name + ".py" in os.listdir(os.path.join(ROOT, "tests")):
                continue  # a test FILE, e.g. `test_abi.py`
            ok = any(t.startswith(name) for t in tests) if star or name.endswith("_") else name in tests
            if not ok:
                missing.append((doc, name))
    assert not missing, missing


def test_every_cited_profile_and_tool_exists():
    missing = []
    for doc in DOCS:
        txt = _text(doc)
        for path in set(re.findall(r"`((?:profiles|tools|tests|include|oracle|rust|raytracer-weekend_b200)/[\w./*{},-]+)`", txt)):
            if "…" in path or path.startswith("oracle/_ref"):   # (oracle/_ref: documented as absent — no Rust toolchain)
                continue
            # expand every {a,b} group, then glob
            pats, done = [path], False
            while not done:
                done, nxt = True, []
                for q in pats:
                    m = re.search(r"\{([^{}]*)\}", q)
                    if m:
                        done = False
                        nxt += [q[:m.start()] + alt + q[m.end():] for alt in m.group(1).split(",")]
                    else:
                        nxt.append(q)
                pats = nxt
            for pat in pats:
                pat = pat.rstrip(".,")
                if not glob.glob(os.path.join(ROOT, pat)) and not glob.glob(os.path.join(ROOT, pat + "*")):
                    missing.append((doc, pat))
    assert not missing, missing


def test_every_documented_switch_is_read_by_the_code():
    src = ""
    for pat in ("raytracer-weekend_b200/csrc/*.cu", "raytracer-weekend_b200/csrc/*.cuh", "raytracer-weekend_b200/*.py",
                "raytracer-weekend_b200/host/*.cpp", "bench.py", "tools/*.py", "tools/*.sh"):
        for p in glob.glob(os.path.join(ROOT, pat)):
            src += open(p).read()
    table = _text("INTEGRATION.md").split("## 6. Run-time switches")[1]
    missing = [v for v in set(re.findall(r"`(RTW_[A-Z0-9_]+)", table)) if v not in src]
    assert not missing, missing
