"""Import alias for the hyphenated package directory `fcd-gan-pytorch_b200/`.

`import fcdgan_b200` resolves sub-modules from `fcd-gan-pytorch_b200/` (a directory name Python cannot
import directly), so user code reads `from fcdgan_b200 import Generator, Segmentor, ...`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fcd-gan-pytorch_b200")
__path__.insert(0, _real)

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
