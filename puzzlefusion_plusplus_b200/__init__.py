"""Import alias for the product package.

The product package directory is ``puzzlefusion-plusplus_b200/`` (the layout the
build contract names); a hyphen is not importable, so this stub re-points
``__path__`` there and executes its ``__init__``.  ``import
puzzlefusion_plusplus_b200`` therefore IS the hyphenated package.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "puzzlefusion-plusplus_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
