"""Import alias: `import lz4java_b200` loads the package that lives in ./lz4-java_b200/
(the hyphen in the directory name is not importable as-is)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lz4-java_b200")]
__package__ = __name__
_init = _os.path.join(__path__[0], "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _os, _f, _init
