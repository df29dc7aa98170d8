"""Importable alias of the `hh-suite_b200/` package directory (a hyphen is not a valid module name)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hh-suite_b200")]
_init = _os.path.join(__path__[0], "__init__.py")
exec(compile(open(_init).read(), _init, "exec"))
