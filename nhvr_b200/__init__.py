"""Importable alias of the product package.

The package directory is named ``neural-human-video-rendering_b200`` (not a valid Python identifier);
this shim makes it importable as ``nhvr_b200`` by pointing ``__path__`` at it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "neural-human-video-rendering_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
