"""Import alias: ``fire_b200`` -> ``face-identification-in-real-time-environments-fire_b200/``.

The build contract fixes the package directory name, which contains hyphens and therefore
cannot be imported directly; this stub points the package path at it and runs its __init__.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "face-identification-in-real-time-environments-fire_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
