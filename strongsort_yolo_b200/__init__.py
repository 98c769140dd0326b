"""Import shim: the package sources live in ``strongsort-yolo_b200/`` (the
directory name the project layout prescribes, which is not a valid Python
identifier).  ``import strongsort_yolo_b200`` resolves to that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "strongsort-yolo_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
