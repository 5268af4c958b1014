"""Import alias for the package directory ``clip-based-cross-modal-hashing_b200/`` (the directory name the
project layout prescribes is not a valid Python identifier).  ``import cmh_b200`` executes that directory's
``__init__.py`` with this module's ``__path__`` pointed at it, so ``cmh_b200.calc_utils`` etc. resolve there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "clip-based-cross-modal-hashing_b200")
__path__ = [_PKG_DIR]
_init = _os.path.join(_PKG_DIR, "__init__.py")
with open(_init, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
