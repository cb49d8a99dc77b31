"""Import shim: the product package lives in the directory `svt-av1-psy_b200/` (the name the build
contract asks for, which is not a legal Python identifier).  This shim makes it importable as
`svt_av1_psy_b200` by pointing the package path at that directory and executing its __init__."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "svt-av1-psy_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
