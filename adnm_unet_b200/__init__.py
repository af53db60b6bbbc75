"""Importable alias of the product package, whose directory is named `adnm-unet_b200/` (not a valid Python
identifier).  `import adnm_unet_b200` executes `adnm-unet_b200/__init__.py` with sub-modules resolved there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "adnm-unet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
