"""Import shim: the product package lives in the directory ``gym-duckietown_b200/`` (a name Python
cannot import directly because of the hyphen).  This module makes ``import gym_duckietown_b200``
resolve to that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gym-duckietown_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
