"""`sdfgenfast_b200` — the repository-name alias of the `sdfgen_b200` package.

The product package is `sdfgen_b200` (it mirrors the reference's Python module `sdfgen`, python/sdfgen.py:47-265,
and owns `csrc/` and `libsdfb.so`).  This name exists so that `import sdfgenfast_b200` and
`sdfgenfast_b200.generate_sdf(...)` work as well; both names are the SAME module objects (no second copy of the
library handle, no second code path)."""
import sys as _sys

import sdfgen_b200 as _pkg
from sdfgen_b200 import *  # noqa: F401,F403
from sdfgen_b200 import _lib, dist, mesh_io, meshes  # noqa: F401

__all__ = getattr(_pkg, "__all__", [n for n in dir(_pkg) if not n.startswith("_")])
for _name in ("_lib", "dist", "mesh_io", "meshes"):
    _sys.modules[__name__ + "." + _name] = _sys.modules["sdfgen_b200." + _name]
for _name in dir(_pkg):
    if not _name.startswith("__"):
        globals().setdefault(_name, getattr(_pkg, _name))
