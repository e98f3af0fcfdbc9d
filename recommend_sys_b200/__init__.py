"""Import shim: the product package lives in `recommend-sys_b200/` (hyphenated, as the layout
contract names it); this makes it importable as `recommend_sys_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "recommend-sys_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
