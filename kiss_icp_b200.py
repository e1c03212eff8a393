"""Import shim: the package directory is named ``kiss-icp_b200`` (not a valid Python identifier);
this module makes it importable as ``kiss_icp_b200`` (and ``kiss_icp_b200.<submodule>``)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "kiss-icp_b200")]
__package__ = __name__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f, _os
