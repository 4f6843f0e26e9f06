"""Import shim: the package directory is named `raytracer-odin_b200/` (a dash is not importable),
so this module turns itself into that package: `import raytracer_odin_b200` and
`from raytracer_odin_b200 import api` resolve into `raytracer-odin_b200/`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "raytracer-odin_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
