"""Import alias for the package directory `medical-cross-modality-domain-adaptation_b200/` (whose name
is not a valid Python identifier): `import pnp_b200` resolves every sub-module from that directory."""
import os as _os

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "medical-cross-modality-domain-adaptation_b200")
__path__ = [_pkg]
with open(_os.path.join(_pkg, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg, "__init__.py"), "exec"))
