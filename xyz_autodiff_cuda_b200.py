"""Import alias: `import xyz_autodiff_cuda_b200` loads the package directory `xyz-autodiff-cuda_b200/`
(the directory name required by the repo layout is not a valid Python identifier)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "xyz-autodiff-cuda_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
