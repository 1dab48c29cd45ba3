"""Import shim: the product package lives in ``raytracer-weekend_b200/`` (hyphenated, as the layout
contract names it), which Python cannot import by name.  This package forwards to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "raytracer-weekend_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
