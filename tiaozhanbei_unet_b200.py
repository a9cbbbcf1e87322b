"""Import shim: makes the package directory ``tiaozhanbei-unet_b200/`` (a name Python cannot
import directly because of the hyphen) importable as ``tiaozhanbei_unet_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "tiaozhanbei-unet_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f, _os
