"""Importable alias of the `show-attend-and-tell_b200/` package (a directory name with
hyphens cannot be imported directly).  `import sat_b200` loads that package in place."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "show-attend-and-tell_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
